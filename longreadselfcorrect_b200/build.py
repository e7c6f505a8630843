"""In-tree build of libpbsc.so (CUDA kernels + C ABI) and the `pbcorrect` host binary for sm_100a."""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libpbsc.so")
BIN = os.path.join(PKG, "pbcorrect")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -fmad=false: the reference's float/double pruning rules must not be contracted into FMAs
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false", "-DPBSC_FUSED_UPDATE",
              "-Xcompiler", "-fPIC,-O3,-ffp-contract=off", "-ccbin", "/usr/bin/g++"]
CU_SOURCES = ["pbsc_index.cu", "pbsc_seed.cu", "pbsc_extend.cu", "pbsc_extend_thread.cu", "pbsc_dp.cu", "pbsc_pipeline.cu", "pbsc_store.cu", "pbsc_build.cu"]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_lib(force: bool = False, verbose: bool = False, extra: list[str] | None = None) -> str:
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(PKG, "..", "include", "pbsc.h"))
    if not force and not _stale(LIB, deps):
        return LIB
    objs = []
    procs = []
    for s in srcs:
        o = s[:-3] + ".o"
        objs.append(o)
        cmd = [NVCC] + NVCC_FLAGS + (extra or []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart", "-ccbin", "/usr/bin/g++"]
    subprocess.run(cmd, check=True)
    return LIB


def build_variant(name: str, extra: list[str], force: bool = False) -> str:
    """An alternative build of the library for experiments (e.g. build_variant("fused", ["-DPBSC_FUSED_UPDATE"])): objects go
    to csrc/_variant_<name>/, the library to libpbsc_<name>.so; select it with PBSC_LIB=<path> (api.py).  The default
    library is not touched."""
    odir = os.path.join(CSRC, f"_variant_{name}")
    os.makedirs(odir, exist_ok=True)
    lib = os.path.join(PKG, f"libpbsc_{name}.so")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))] + [os.path.join(PKG, "..", "include", "pbsc.h")]
    if not force and not _stale(lib, deps):
        return lib
    procs, objs = [], []
    for s in CU_SOURCES:
        o = os.path.join(odir, s[:-3] + ".o")
        objs.append(o)
        cmd = [NVCC] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    subprocess.run([NVCC, "-shared", "-o", lib] + objs + ["-lcudart", "-ccbin", "/usr/bin/g++"], check=True)
    return lib


def build_cli(force: bool = False) -> str:
    src = os.path.join(CSRC, "pbcorrect_main.cpp")
    if not os.path.exists(src):
        return ""
    if not force and not _stale(BIN, [src, LIB, os.path.join(PKG, "..", "include", "pbsc.h")]):
        return BIN
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", src, "-o", BIN, "-I", os.path.join(PKG, "..", "include"),
           "-L", PKG, "-lpbsc", "-Wl,-rpath,$ORIGIN", "-lpthread", "-lz"]
    subprocess.run(cmd, check=True)
    return BIN


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":   # python -m longreadselfcorrect_b200.build --variant fused -DPBSC_FUSED_UPDATE
        print(build_variant(sys.argv[2], sys.argv[3:]))
        sys.exit(0)
    build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_cli(force="--force" in sys.argv)
    print(LIB)
