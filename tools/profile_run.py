#!/usr/bin/env python3
"""One pass of the hot path on a small synthetic workload, for ncu (see profiles/README.md).
    python tools/profile_run.py [workload] [k0] [reps] [dp]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bench  # noqa: E402
from longreadselfcorrect_b200 import api, bwt_build  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "mini"]
k0 = int(sys.argv[2]) if len(sys.argv) > 2 else 13
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dp = len(sys.argv) > 4 and sys.argv[4] == "dp"
codes, off = bench.make_data(wl)
n = off.size - 1
runs = {}
for ext, rev in (("bwt", False), ("rbwt", True)):
    b = bwt_build.bwt_symbols(codes, off, reverse=rev, device="cuda:0")
    runs[ext] = (bwt_build.run_length_bytes(b), int(b.numel()), n)
idx = api.Index.from_runs(runs["bwt"][0], runs["bwt"][1], n, runs["rbwt"][0], runs["rbwt"][1], n)
if k0:
    idx.build_prefix_table(k0)
p = api.Params.make(coverage=wl["c"], genome=wl["g"], no_dp=not dp)
batch = api.Batch(idx, p, packed=bench.packed_ascii(codes, off))
for _ in range(reps):
    ms = batch.run()
    print("run ms", ms, api.last_timing())
