#!/usr/bin/env python3
"""Where the end-to-end call (pbsc_correct_batch on host buffers) spends host wall time.
    PBSC_TRACE=1 python tools/e2e_trace.py [workload] [dp]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from longreadselfcorrect_b200 import api, bwt_build  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg1"]
dp = len(sys.argv) > 2 and sys.argv[2] == "dp"
codes, off = bench.make_data(wl)
n = off.size - 1
runs = {}
for ext, rev in (("bwt", False), ("rbwt", True)):
    b = bwt_build.bwt_symbols(codes, off, reverse=rev, device="cuda:0")
    runs[ext] = (bwt_build.run_length_bytes(b), int(b.numel()), n)
idx = api.Index.from_runs(runs["bwt"][0], runs["bwt"][1], n, runs["rbwt"][0], runs["rbwt"][1], n)
idx.build_prefix_table(13)
p = api.Params.make(coverage=wl["c"], genome=wl["g"], no_dp=not dp)
packed = bench.packed_ascii(codes, off)
packed = (api.pinned_copy(packed[0]), api.pinned_copy(packed[1]))
for i in range(4):
    t = time.perf_counter()
    idx.correct_reads(p, packed=packed, pinned_out=True)
    print(f"correct_reads call {i}: {1000 * (time.perf_counter() - t):.1f} ms wall", api.last_timing(), flush=True)
