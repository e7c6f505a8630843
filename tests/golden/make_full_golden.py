#!/usr/bin/env python3
"""Whole-output goldens at BASELINE sizes: per-read digests of what the UNMODIFIED reference writes.

    python tests/golden/make_full_golden.py cfg2                 # every read of config 2 (CPU only, ~25 min on 8 cores)
    python tests/golden/make_full_golden.py cfg3 --sample 0.05   # a seeded 5 % NON-PREFIX sample of config 3 (run on the GPU box:
                                                                 # the 1.24 Gbp BWT is built with bwt_build on the GPU)

The reads and the index are the ones bench.py uses (same seeds, same bwt_build files), the binary is oracle/_ref/stride
(built by oracle/build_ref.py from /root/reference, unmodified), the command is `stride pbcorrect -t <cores> -p idx -o out
-c C -g G reads.fa`.  For every read id the file stores one 8-byte record: the first 7 bytes of
sha256(b">" + id + b"\n" + sequence + b"\n") of the read's record in correct.fa or discard.fa, and one byte that says which
file it was in (1 = correct.fa, 0 = discard.fa).  `longreadselfcorrect_b200.parity` computes the same digests from a GPU
result, so bench.py and the GPU tests can check EVERY read of the timed run, not a prefix sample.

Output: tests/golden/<workload>.read_sha.npz  (ids: int32 read indices, digest: uint8 [n, 8], meta: json string)
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def record_digest(rid: bytes, seq: bytes, in_correct: bool) -> bytes:
    return hashlib.sha256(b">" + rid + b"\n" + seq + b"\n").digest()[:7] + (b"\x01" if in_correct else b"\x00")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--sample", type=float, default=1.0, help="fraction of reads (seeded, spread over the whole set)")
    ap.add_argument("--sample-seed", type=int, default=20261018)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--device", default=None, help="torch device for bwt_build (default: cuda if available)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--nodp", action="store_true", help="run the reference with --nodp (golden name <workload>_nodp)")
    args = ap.parse_args()
    import bench
    from longreadselfcorrect_b200 import bwt_build, synth

    wl = bench.WORKLOADS[args.workload]
    codes, off = bench.make_data(wl)
    n = off.size - 1
    if args.sample < 1.0:
        rng = np.random.Generator(np.random.PCG64(args.sample_seed))
        ids = np.sort(rng.choice(n, size=max(1, int(round(n * args.sample))), replace=False)).astype(np.int32)
    else:
        ids = np.arange(n, dtype=np.int32)
    t0 = time.time()
    with tempfile.TemporaryDirectory(dir=os.environ.get("TMPDIR")) as d:
        prefix = os.path.join(d, "idx")
        bwt_build.build_index_files(prefix, codes, off, device=args.device)
        print(f"index files written in {time.time() - t0:.0f}s", file=sys.stderr, flush=True)
        fa = os.path.join(d, "reads.fa")
        letters = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
        with open(fa, "wb") as f:
            for i in ids:
                f.write(b">r%d\n" % i)
                f.write(letters[off[i]:off[i + 1]].tobytes())
                f.write(b"\n")
        t1 = time.time()
        out = os.path.join(d, "out")
        r = subprocess.run([bench.REF_STRIDE, "pbcorrect", "-t", str(args.threads), "-p", prefix, "-o", out, "-c", str(wl["c"]), "-g", str(wl["g"])] + (["--nodp"] if args.nodp else []) + [fa],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        if r.returncode != 0:
            raise SystemExit("reference failed: " + r.stderr[-2000:])
        secs = time.time() - t1
        print(f"reference: {ids.size} reads in {secs:.0f}s with {args.threads} threads", file=sys.stderr, flush=True)
        dig = {}
        for fn, flag in (("correct.fa", True), ("discard.fa", False)):
            name = None
            with open(os.path.join(out, fn), "rb") as f:
                for line in f:
                    if line.startswith(b">"):
                        name = line[1:].strip()
                    else:
                        dig[int(name[1:])] = record_digest(name, line.rstrip(b"\n"), flag)
        summary = r.stdout
    missing = [int(i) for i in ids if int(i) not in dig]
    if missing:
        raise SystemExit(f"{len(missing)} reads missing from the reference output, e.g. {missing[:5]}")
    digest = np.frombuffer(b"".join(dig[int(i)] for i in ids), dtype=np.uint8).reshape(-1, 8)
    mbp = float(sum(int(off[i + 1] - off[i]) for i in ids)) / 1e6
    meta = {"workload": args.workload + ("_nodp" if args.nodp else ""), "desc": wl["desc"], "reads_total": int(n), "reads_in_file": int(ids.size), "sample": args.sample,
            "sample_seed": args.sample_seed if args.sample < 1.0 else None, "mbp": mbp, "reference_seconds": secs, "reference_threads": args.threads,
            "reference_mbp_per_s": mbp / secs, "command": f"stride pbcorrect -t {args.threads} -p idx -o out -c {wl['c']} -g {wl['g']}{' --nodp' if args.nodp else ''} reads.fa",
            "record": "sha256(b'>' + id + b'\\n' + seq + b'\\n')[:7] + (1 if in correct.fa else 0)", "summary": summary}
    path = args.out or os.path.join(ROOT, "tests", "golden", f"{args.workload}{'_nodp' if args.nodp else ''}.read_sha.npz")
    np.savez_compressed(path, ids=ids, digest=digest, meta=np.array(json.dumps(meta)))
    print(path, json.dumps({k: v for k, v in meta.items() if k != "summary"}))


if __name__ == "__main__":
    main()
