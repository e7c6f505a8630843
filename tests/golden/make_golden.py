#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ with the UNMODIFIED reference binary.

Run in the build container (needs oracle/_ref/stride and oracle/_ref/fm_dump, built from
/root/reference by oracle/build_ref.py):

    python tests/golden/make_golden.py

Fixture `tiny`: 20 kb uniform genome (seed 11) with one injected repeat family, 30x CLR-like reads
(mean 1.5 kb, seed 111), indexed by `stride index`, corrected by
`stride pbcorrect -t 1 -c 30 -g 5 --nodp --debugseed`.  Files:
    tiny.reads.fa, tiny.bwt, tiny.rbwt, tiny.sai        inputs (index built by the reference)
    tiny.correct.fa, tiny.discard.fa, tiny.threshold-table, tiny.summary.txt   reference outputs
    tiny.seeds.tsv       concatenated seed/<id>.seed dumps ("#id" header lines)
    tiny.ext.tsv         concatenated extend/<id>.ext dumps (failed walks + failure code)
    tiny.fm_queries.txt / tiny.fm_bwt.txt / tiny.fm_rbwt.txt   findInterval inputs and reference (lower upper)
Fixture `tiny100`: the same reads corrected with `-c 100 -g 10` (different pool, offsets and thresholds).

    python tests/golden/make_golden.py dp

adds the DEFAULT-option fixtures (DP / multiple-alignment fallback on) from the committed tiny index and reads:
    tiny.dp.correct.fa, tiny.dp.discard.fa, tiny.dp.summary.txt, tiny.dp.dpfail.tsv (extend/<id>.dp rows: pairs where the
    DP fallback failed too), and the same for tiny100.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from longreadselfcorrect_b200 import synth  # noqa: E402

STRIDE = os.path.join(ROOT, "oracle", "_ref", "stride")
FMDUMP = os.path.join(ROOT, "oracle", "_ref", "fm_dump")


def run(cmd, **kw):
    return subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, **kw)


def main():
    tmp = tempfile.mkdtemp(prefix="pbsc_golden_")
    g = synth.make_genome(20000, 11, repeat_families=1)
    codes, off = synth.simulate_reads(g, 30, 1500, 111, min_len=300)
    reads = os.path.join(tmp, "tiny.reads.fa")
    synth.write_fasta(reads, codes, off)
    run([STRIDE, "index", "-t", "2", "-p", os.path.join(tmp, "tiny"), reads], cwd=tmp)
    for ext in ("bwt", "rbwt", "sai"):
        shutil.copy(os.path.join(tmp, "tiny." + ext), os.path.join(HERE, "tiny." + ext))
    shutil.copy(reads, os.path.join(HERE, "tiny.reads.fa"))
    for name, opts in (("tiny", ["-c", "30", "-g", "5"]), ("tiny100", ["-c", "100", "-g", "10"])):
        out = os.path.join(tmp, name + "_out")
        r = run([STRIDE, "pbcorrect", "-t", "1", "-p", os.path.join(tmp, "tiny"), "-o", out] + opts + ["--nodp", "--debugseed", reads], cwd=tmp)
        shutil.copy(os.path.join(out, "correct.fa"), os.path.join(HERE, name + ".correct.fa"))
        shutil.copy(os.path.join(out, "discard.fa"), os.path.join(HERE, name + ".discard.fa"))
        shutil.copy(os.path.join(out, "threshold-table"), os.path.join(HERE, name + ".threshold-table"))
        summary = "\n".join(l for l in r.stdout.splitlines() if not l.startswith("Time of"))
        open(os.path.join(HERE, name + ".summary.txt"), "w").write(summary + "\n")
        n = off.size - 1
        with open(os.path.join(HERE, name + ".seeds.tsv"), "w") as f:
            for i in range(n):
                p = os.path.join(out, "seed", f"r{i}.seed")
                f.write(f"#r{i}\n")
                if os.path.exists(p):
                    f.write(open(p).read())
        with open(os.path.join(HERE, name + ".ext.tsv"), "w") as f:
            for i in range(n):
                p = os.path.join(out, "extend", f"r{i}.ext")
                f.write(f"#r{i}\n")
                if os.path.exists(p):
                    f.write(open(p).read())
    # findInterval vectors: k-mers sampled from the reads (hits), their mutations and random k-mers (early exit)
    rng = np.random.Generator(np.random.PCG64(7))
    strs = synth.read_strings(codes, off)
    qs = []
    for _ in range(1500):
        s = strs[int(rng.integers(0, len(strs)))]
        k = int(rng.integers(1, 52))
        if len(s) <= k:
            continue
        p = int(rng.integers(0, len(s) - k))
        w = s[p:p + k]
        if rng.random() < 0.3:
            j = int(rng.integers(0, k))
            w = w[:j] + "ACGT"[int(rng.integers(0, 4))] + w[j + 1:]
        qs.append(w)
    for _ in range(500):
        k = int(rng.integers(1, 40))
        qs.append("".join("ACGT"[int(x)] for x in rng.integers(0, 4, size=k)))
    qf = os.path.join(HERE, "tiny.fm_queries.txt")
    open(qf, "w").write("\n".join(qs) + "\n")
    for which in ("bwt", "rbwt"):
        r = run([FMDUMP, os.path.join(HERE, "tiny." + which), qf])
        open(os.path.join(HERE, f"tiny.fm_{which}.txt"), "w").write(r.stdout)
    shutil.rmtree(tmp)
    print("golden fixtures written to", HERE)


def dp_unit_vectors():
    """Direct vectors for Overlapper::extendMatch and MultipleAlignment: inputs in dp_units.txt, the reference's answers
    (oracle/_ref/dp_dump, linked against the reference's own objects) in dp_units.ref.txt."""
    rng = np.random.Generator(np.random.PCG64(2025))
    acgt = "ACGT"

    def rand_seq(n):
        return "".join(acgt[int(x)] for x in rng.integers(0, 4, size=n))

    def mutate(s, rate):
        out = []
        for ch in s:
            u = rng.random()
            if u < rate * 0.3:
                continue                                  # deletion
            if u < rate * 0.45:
                out.append(acgt[int(rng.integers(0, 4))])  # substitution
            else:
                out.append(ch)
            if rng.random() < rate * 0.55:
                out.append(out[-1] if rng.random() < 0.5 else acgt[int(rng.integers(0, 4))])   # insertion (often a homopolymer)
        return "".join(out) or "A"

    lines = []
    for i in range(360):
        n = int(rng.integers(20, 420))
        q = rand_seq(n)
        if i % 7 == 0:   # low-complexity stretches exercise the homopolymer tie-breaks
            p0 = int(rng.integers(0, max(1, n - 12)))
            q = q[:p0] + acgt[int(rng.integers(0, 4))] * 12 + q[p0 + 12:]
        k = int(rng.integers(9, 20))
        if n <= k + 2:
            continue
        mode = i % 4
        if mode == 0:     # read starts with the query's first k-mer and runs on (forward seed)
            m = q[:k] + mutate(q[k:], 0.13) + rand_seq(int(rng.integers(0, 60)))
            lines.append(f"A {q} {m} 0 0")
        elif mode == 1:   # read ends with the query's last k-mer (reverse seed)
            m = rand_seq(int(rng.integers(0, 60))) + mutate(q[:-k], 0.13) + q[-k:]
            lines.append(f"A {q} {m} {len(q) - k} {len(m) - k}")
        elif mode == 2:   # short read: the band leaves the matrix
            m = q[:k] + mutate(q[k:k + int(rng.integers(1, 40))], 0.2)
            lines.append(f"A {q} {m} 0 0")
        else:             # unrelated read sharing only the seed
            m = q[:k] + rand_seq(int(rng.integers(5, 300)))
            lines.append(f"A {q} {m} 0 0")
    for i in range(60):
        n = int(rng.integers(60, 300))
        q = rand_seq(n)
        k = 13
        rows = []
        for r in range(int(rng.integers(3, 25))):
            if r % 2 == 0:
                m = q[:k] + mutate(q[k:], 0.13)
                rows.append(f"{m} 0 0")
            else:
                m = mutate(q[:-k], 0.13) + q[-k:]
                rows.append(f"{m} {len(q) - k} {len(m) - k}")
        lines.append(f"M {q} {int(rng.integers(3, 20))} {len(rows)}")
        lines.extend(rows)
    path = os.path.join(HERE, "dp_units.txt")
    open(path, "w").write("\n".join(lines) + "\n")
    r = run([os.path.join(ROOT, "oracle", "_ref", "dp_dump"), path])
    open(os.path.join(HERE, "dp_units.ref.txt"), "w").write(r.stdout)
    print("DP unit vectors:", len(r.stdout.splitlines()), "records")


def dp_fixtures():
    tmp = tempfile.mkdtemp(prefix="pbsc_golden_dp_")
    reads = os.path.join(HERE, "tiny.reads.fa")
    n = sum(1 for l in open(reads) if l.startswith(">"))
    for name, opts in (("tiny", ["-c", "30", "-g", "5"]), ("tiny100", ["-c", "100", "-g", "10"])):
        out = os.path.join(tmp, name + "_out")
        r = run([STRIDE, "pbcorrect", "-t", "1", "-p", os.path.join(HERE, "tiny"), "-o", out] + opts + ["--debugseed", reads], cwd=tmp)
        shutil.copy(os.path.join(out, "correct.fa"), os.path.join(HERE, name + ".dp.correct.fa"))
        shutil.copy(os.path.join(out, "discard.fa"), os.path.join(HERE, name + ".dp.discard.fa"))
        summary = "\n".join(l for l in r.stdout.splitlines() if not l.startswith("Time of"))
        open(os.path.join(HERE, name + ".dp.summary.txt"), "w").write(summary + "\n")
        with open(os.path.join(HERE, name + ".dp.dpfail.tsv"), "w") as f:
            for i in range(n):
                p = os.path.join(out, "extend", f"r{i}.dp")
                f.write(f"#r{i}\n")
                if os.path.exists(p):
                    f.write(open(p).read())
    shutil.rmtree(tmp)
    print("DP golden fixtures written to", HERE)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "dp":
        dp_fixtures()
        dp_unit_vectors()
    elif len(sys.argv) > 1 and sys.argv[1] == "dpunits":
        dp_unit_vectors()
    else:
        main()
