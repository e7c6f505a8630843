#!/usr/bin/env python3
"""Build the UNMODIFIED reference (`stride`) from /root/reference into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product; the
product (longreadselfcorrect_b200/) never imports, links or executes it.

Recipe = SURVEY.md section 8c: the reference's autotools build cannot run in
this image (no autoconf/automake, no google-sparsehash), so we compile exactly
the sources named in the reference's eleven */Makefile.am files, where they lie
under /root/reference, with a hand-written config.h, a sparsehash shim and a
force-included compat header (oracle/refbuild/).  Flags follow configure.ac:71-76
minus -static: `-O3 -std=c++11 -fpermissive -fopenmp`, no -march (so no FMA
contraction: plain IEEE float/double, which the GPU path must match).

Outputs (git-ignored, but shipped to the GPU box by gpurun):
  oracle/_ref/stride          the reference binary
  oracle/_ref/fm_dump         tiny driver linked against the reference objects;
                              prints BWTAlgorithms::findInterval lower/upper
  oracle/_ref/dp_dump         the same for Overlapper::extendMatch and MultipleAlignment (DP fallback)
  oracle/_ref/obj/*.o         objects
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PBSC_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
OBJ = os.path.join(OUT, "obj")
SUBDIRS = ["Util", "SQG", "Bigraph", "Algorithm", "StringGraph", "Concurrency",
           "SuffixTools", "FMIndexWalk", "PacBio", "Thirdparty", "StriDe"]


def sources():
    """(dir, relative source) for every .cpp/.c/.C named in a *_SOURCES list."""
    seen, out = set(), []
    for d in SUBDIRS:
        am = open(os.path.join(REF, d, "Makefile.am")).read().replace("\\\n", " ")
        for m in re.finditer(r"^\s*\w+_SOURCES\s*=(.*)$", am, re.M):
            for tok in m.group(1).split():
                if tok.endswith((".cpp", ".c", ".C")):
                    key = os.path.join(d, tok)
                    if key not in seen and os.path.exists(os.path.join(REF, key)):
                        seen.add(key)
                        out.append(key)
    return out


def main():
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} not present: keeping prebuilt oracle/_ref as is")
        return 0
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    incs = " ".join(f"-I{os.path.join(REF, d)}" for d in SUBDIRS)
    incs += f" -I{REF}/Thirdparty/ropebwt2 -I{REF}/Thirdparty/rollinghash"
    incs += f" -I{HERE}/refbuild -I{HERE}/refbuild/shim"
    cxx = ("g++ -O3 -std=c++11 -fpermissive -fopenmp -w -DHAVE_CONFIG_H "
           f"-include {HERE}/refbuild/compat.h {incs}")
    cc = f"gcc -O3 -fopenmp -w -DHAVE_CONFIG_H {incs}"
    objs, rules = [], []
    for s in srcs:
        o = os.path.join(OBJ, s.replace("/", "_").rsplit(".", 1)[0] + ".o")
        objs.append(o)
        comp = cc if s.endswith(".c") else cxx
        rules.append(f"{o}: {os.path.join(REF, s)}\n\t{comp} -c $< -o $@\n")
    main_o = [o for o in objs if o.endswith("StriDe_StriDe.o")]
    lib_o = [o for o in objs if o not in main_o]
    drv = os.path.join(HERE, "refbuild", "fm_dump.cpp")
    drv2 = os.path.join(HERE, "refbuild", "dp_dump.cpp")
    mk = os.path.join(OUT, "Makefile")
    with open(mk, "w") as f:
        f.write(f"all: {OUT}/stride {OUT}/fm_dump {OUT}/dp_dump\n")
        f.write(f"{OUT}/stride: {' '.join(objs)}\n\tg++ -fopenmp -pthread $^ -lz -o $@\n")
        f.write(f"{OUT}/fm_dump: {drv} {' '.join(lib_o)}\n\t{cxx} $^ -pthread -lz -o $@\n")
        f.write(f"{OUT}/dp_dump: {drv2} {' '.join(lib_o)}\n\t{cxx} $^ -pthread -lz -o $@\n")
        f.write("\n".join(rules))
    r = subprocess.run(["make", "-f", mk, f"-j{os.cpu_count() or 4}", "-s"], cwd=OUT)
    return r.returncode


if __name__ == "__main__":
    sys.exit(main())
