#!/usr/bin/env python3
"""Golden outputs of `stride pbcorrect --onlyseed -b BARCODE` (seed validation against alignment barcodes, PacBio/BCode.cpp;
PacBioSelfCorrectionProcess.cpp:265-335,372-380) from the UNMODIFIED reference binary (oracle/_ref/stride).

    python tests/golden/make_onlyseed_golden.py

There is no aligner here, so the barcode file is synthetic: for the first reads of tiny.reads.fa one or two blocks of a few hundred
bases with a random code string (two hex digits per base: mostly '0', a few insertion marks '1' in the even places, a few deletion
marks in the odd ones) and a random strand.  What is pinned is the reference's arithmetic on such a file, not biology.
Writes tests/golden/tiny.barcode.txt, tiny.onlyseed.total.seed (DIR/total.seed) and tiny.onlyseed.stdout, and, from
`stride kmercheck -c 30 -l 15 -u 23 -s 4` on the same barcode file, tiny.kmercheck.total.box and tiny.kmercheck.value.box."""
import os
import random
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
STRIDE = os.path.join(ROOT, "oracle", "_ref", "stride")
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import read_fasta  # noqa: E402


def main():
    random.seed(20261019)
    recs = read_fasta(os.path.join(HERE, "tiny.reads.fa"))
    lines = []
    for name, seq in recs[:150]:
        L = len(seq)
        cuts = []
        if L > 900:
            a = random.randrange(0, L // 2 - 300)
            cuts.append((a, a + random.randrange(150, 300)))
            a2 = random.randrange(L // 2, L - 300)
            cuts.append((a2, min(L, a2 + random.randrange(150, 300))))
        else:
            cuts.append((0, L))
        for a, b in cuts:
            code = []
            for _ in range(b - a):
                u, v = random.random(), random.random()
                up = "1" if u < 0.03 else "0"
                lo = "0" if v < 0.96 else (random.choice("1248") if v < 0.99 else random.choice("3569ac"))
                code.append(up + lo)
            lines.append(f"{name}\t{a}\t{b}\tref\t{a + 1000}\t{b + 1000}\t{''.join(code)}\t{'True' if random.random() < 0.5 else 'False'}\t0")
    bc = os.path.join(HERE, "tiny.barcode.txt")
    open(bc, "w").write("\n".join(lines) + "\n")
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run([STRIDE, "pbcorrect", "-t", "1", "-p", os.path.join(HERE, "tiny"), "-o", os.path.join(d, "out"), "-c", "30", "-g", "5", "--onlyseed", "-b", bc,
                            os.path.join(HERE, "tiny.reads.fa")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, check=True)
        shutil.copy(os.path.join(d, "out", "total.seed"), os.path.join(HERE, "tiny.onlyseed.total.seed"))
        open(os.path.join(HERE, "tiny.onlyseed.stdout"), "w").write(r.stdout)
        subprocess.run([STRIDE, "kmercheck", "-t", "1", "-c", "30", "-p", os.path.join(HERE, "tiny"), "-o", os.path.join(d, "kc"), "-b", bc, "-l", "15", "-u", "23", "-s", "4",
                        os.path.join(HERE, "tiny.reads.fa")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, check=True)
        for name in ("total.box", "value.box"):
            shutil.copy(os.path.join(d, "kc", name), os.path.join(HERE, "tiny.kmercheck." + name))
        print(r.stdout.strip(), "|", len(open(os.path.join(d, "out", "total.seed")).read().splitlines()), "reads with invalid seeds |",
              sorted(os.listdir(os.path.join(d, "out"))))


if __name__ == "__main__":
    main()
