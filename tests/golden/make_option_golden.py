#!/usr/bin/env python3
"""Goldens for options the other fixtures leave at their defaults, written by the UNMODIFIED reference binary
(oracle/_ref/stride) on the committed tiny index and reads:

    python tests/golden/make_option_golden.py

tests/golden/tiny.options.json: for every variant the option list, the sha256 of the reference's correct.fa and
discard.fa, and the counters of its stdout summary.  Variants: -k/-u/-r (the `adjust` branch of
StriDe/PacBioSelfCorrection.cpp:195-206), -i 7, -s 15, -e 0.2, -l 16, -m 2, and FASTQ input (Util/SeqReader.cpp:71-99)."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
STRIDE = os.path.join(ROOT, "oracle", "_ref", "stride")
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import read_fasta  # noqa: E402

VARIANTS = {
    "adjust_k17_u2_r-2": ["-c", "30", "-g", "5", "-k", "17", "-u", "2", "-r", "-2"],
    "adjust_k21": ["-c", "30", "-g", "10", "-k", "21"],
    "idmer7": ["-c", "30", "-g", "5", "-i", "7"],
    "minkmer15": ["-c", "30", "-g", "5", "-s", "15"],
    "error0.2": ["-c", "30", "-g", "5", "-e", "0.2"],
    "leaves16": ["-c", "30", "-g", "5", "-l", "16"],
    "mode2_nodp": ["-c", "30", "-g", "5", "-m", "2", "--nodp"],
    "fastq": ["-c", "30", "-g", "5"],
}


def write_fastq(path, reads):
    with open(path, "w") as f:
        for i, (rid, seq) in enumerate(reads):
            # quality strings that start with '@' and '>' must not be taken for headers
            q = ("@" if i % 3 == 0 else ">" if i % 3 == 1 else "I") + "I" * (len(seq) - 1)
            f.write(f"@{rid} some comment\n{seq.lower() if i % 5 == 0 else seq}\n+\n{q}\n")


def main():
    reads = read_fasta(os.path.join(HERE, "tiny.reads.fa"))
    out = {}
    with tempfile.TemporaryDirectory() as d:
        fq = os.path.join(d, "tiny.reads.fq")
        write_fastq(fq, reads)
        for name, opts in VARIANTS.items():
            o = os.path.join(d, name)
            src = fq if name == "fastq" else os.path.join(HERE, "tiny.reads.fa")
            r = subprocess.run([STRIDE, "pbcorrect", "-t", "1", "-p", os.path.join(HERE, "tiny"), "-o", o] + opts + [src], check=True,
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            summ = {l.split(":")[0]: l.split(":")[1].split(",")[0].strip() for l in r.stdout.splitlines() if ":" in l and not l.startswith("Time of")}
            out[name] = {"options": opts, "input": "fastq" if name == "fastq" else "fasta",
                         "correct_sha256": hashlib.sha256(open(os.path.join(o, "correct.fa"), "rb").read()).hexdigest(),
                         "discard_sha256": hashlib.sha256(open(os.path.join(o, "discard.fa"), "rb").read()).hexdigest(),
                         "correct_bytes": os.path.getsize(os.path.join(o, "correct.fa")), "summary": summ}
            print(name, out[name]["correct_sha256"][:16], summ.get("FMNum"), summ.get("DPNum"), flush=True)
    json.dump(out, open(os.path.join(HERE, "tiny.options.json"), "w"), indent=1)




def debug_dump_digest(outdir, ids):
    """sha256 per dump-file class of a --debugseed run: the files of all reads concatenated in read order, each preceded by
    '#<id>\\n', or '#<id> missing\\n' when the reference did not create it."""
    res = {}
    for cls, pat in (("seed", "seed/%s.seed"), ("seed_error", "seed/error/%s.seed"), ("log", "extend/%s.log"), ("ext", "extend/%s.ext"), ("dp", "extend/%s.dp")):
        h = hashlib.sha256()
        n = 0
        for rid in ids:
            p = os.path.join(outdir, pat % rid)
            if os.path.exists(p):
                h.update(b"#" + rid.encode() + b"\n")
                h.update(open(p, "rb").read())
                n += 1
            else:
                h.update(b"#" + rid.encode() + b" missing\n")
        res[cls] = {"sha256": h.hexdigest(), "files": n}
    return res


def debugseed_goldens():
    """tests/golden/tiny.debugseed.json: the dump files of `stride pbcorrect --debugseed` with and without --nodp."""
    reads = read_fasta(os.path.join(HERE, "tiny.reads.fa"))
    ids = [r for r, _ in reads]
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for name, opts in (("nodp", ["-c", "30", "-g", "5", "--nodp"]), ("default", ["-c", "30", "-g", "5"])):
            o = os.path.join(d, name)
            subprocess.run([STRIDE, "pbcorrect", "-t", "1", "-p", os.path.join(HERE, "tiny"), "-o", o, "--debugseed"] + opts + [os.path.join(HERE, "tiny.reads.fa")],
                           check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            out[name] = {"options": opts + ["--debugseed"], "dumps": debug_dump_digest(o + "/", ids)}
            print(name, {k: (v["files"], v["sha256"][:12]) for k, v in out[name]["dumps"].items()}, flush=True)
    json.dump(out, open(os.path.join(HERE, "tiny.debugseed.json"), "w"), indent=1)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "debugseed":
        debugseed_goldens()
    else:
        main()
