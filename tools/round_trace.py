#!/usr/bin/env python3
"""Where does a batch's time go?  Runs the first 1/PARTS of a workload's reads as one batch with PBSC_ROUND_TRACE=1 and
PBSC_DP_PROFILE=1 (per-round host wall clock, per-launch walk kernel times, per-kernel DP stage times on stderr).
    python tools/round_trace.py [workload] [parts] [part]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bench  # noqa: E402
from longreadselfcorrect_b200 import api, bwt_build, sharding  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 4
part = int(sys.argv[3]) if len(sys.argv) > 3 else 0
codes, off = bench.make_data(wl)
n = off.size - 1
runs = {}
for ext, rev in (("bwt", False), ("rbwt", True)):
    b = bwt_build.bwt_symbols(codes, off, reverse=rev, device="cuda:0")
    runs[ext] = (bwt_build.run_length_bytes(b), int(b.numel()), n)
idx = api.Index.from_runs(runs["bwt"][0], runs["bwt"][1], n, runs["rbwt"][0], runs["rbwt"][1], n)
idx.build_prefix_table(13)
p = api.Params.make(coverage=wl["c"], genome=wl["g"])
letters, off64 = bench.packed_ascii(codes, off)
b0, b1 = sharding.balanced_ranges(np.diff(off), parts)[part]
o = off64[b0:b1 + 1]
batch = api.Batch(idx, p, packed=(np.ascontiguousarray(letters[int(o[0]):int(o[-1])]), (o - o[0]).astype(np.uint64)))
batch.run()
os.environ["PBSC_ROUND_TRACE"] = "1"
os.environ["PBSC_DP_PROFILE"] = "1"
print(f"---- traced pass: reads [{b0}, {b1}) = {int(o[-1] - o[0]) / 1e6:.1f} Mbp ----", file=sys.stderr, flush=True)
ms = batch.run()
print("run ms", ms, api.last_timing(), file=sys.stderr)
