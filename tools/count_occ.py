#!/usr/bin/env python3
"""Issued index sectors per kernel family on a sample of reads (run with PBSC_LIB=.../libpbsc_count.so, the -DPBSC_COUNT_OCC
build of the same kernels):  python tools/count_occ.py PREFIX READS.fa COVERAGE GENOME K0 DEVICE [--nodp]
Prints one JSON line: {"seed", "setup", "walk", "dp", "other", "walks", "bases"}."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from longreadselfcorrect_b200 import api  # noqa: E402

prefix, fa, cov, genome, k0, dev = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
nodp = "--nodp" in sys.argv[7:]
reads = [l.strip() for l in open(fa) if not l.startswith(">")]
idx = api.Index.load(prefix, device=dev)
if k0:
    idx.build_prefix_table(k0)
p = api.Params.make(coverage=cov, genome=genome, no_dp=nodp)
b = api.Batch(idx, p, reads=reads)
api.occ_counts(reset=True)
b.run()
ok, c = api.occ_counts(reset=True)
assert ok, "PBSC_LIB must point at the -DPBSC_COUNT_OCC build"
c["walks"] = int(api.last_timing()["seed_pairs"])
c["bases"] = int(sum(len(r) for r in reads))
print(json.dumps(c))
