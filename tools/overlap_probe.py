#!/usr/bin/env python3
"""Does running two batches concurrently (two streams, two scratch arenas) fill the tails of the round structure?

    python tools/overlap_probe.py [--workload cfg2] [--parts 4] [--lanes 2] [--reps 2]

The read set is cut into `parts` length-balanced batches.  They are run (kernels only, reads resident) first one after the
other, then by `lanes` host threads at once; the wall time of a whole pass is printed for both."""
import argparse
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--parts", type=int, default=4)
    ap.add_argument("--lanes", type=int, default=2)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--nodp", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    from longreadselfcorrect_b200 import api, bwt_build, sharding

    wl = bench.WORKLOADS[args.workload]
    codes, off = bench.make_data(wl)
    n = off.size - 1
    runs = {}
    for ext, rev in (("bwt", False), ("rbwt", True)):
        b = bwt_build.bwt_symbols(codes, off, reverse=rev, device="cuda:0")
        runs[ext] = (bwt_build.run_length_bytes(b), int(b.numel()), n)
    packed = bench.packed_ascii(codes, off)
    mbp = codes.size / 1e6
    idx = api.Index.from_runs(runs["bwt"][0], runs["bwt"][1], n, runs["rbwt"][0], runs["rbwt"][1], n)
    idx.build_prefix_table(13)
    if hasattr(idx, "set_lanes"):
        idx.set_lanes(args.lanes)
    p = api.Params.make(coverage=wl["c"], genome=wl["g"], no_dp=args.nodp)
    chunks = []
    for b0, b1 in sharding.balanced_ranges(np.diff(off), args.parts):
        o = off[b0:b1 + 1]
        chunks.append((np.ascontiguousarray(packed[0][int(o[0]):int(o[-1])]), (o - o[0]).astype(np.uint64)))
    batches = [api.Batch(idx, p, packed=c) for c in chunks]

    def one_pass(lanes):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if lanes <= 1:
            for b in batches:
                b.run()
        else:
            nxt = [0]
            lock = threading.Lock()

            def work():
                while True:
                    with lock:
                        i = nxt[0]
                        nxt[0] += 1
                    if i >= len(batches):
                        return
                    batches[i].run()
            th = [threading.Thread(target=work) for _ in range(lanes)]
            for t in th:
                t.start()
            for t in th:
                t.join()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3

    for lanes in (1, args.lanes):
        one_pass(lanes)   # warm-up of the arenas of every lane
        ts = [one_pass(lanes) for _ in range(args.reps)]
        ms = min(ts)
        free_b, tot_b = torch.cuda.mem_get_info()
        print(f"{args.workload} parts={args.parts} lanes={lanes}: {ms:8.1f} ms per pass  {mbp / (ms / 1e3):7.1f} Mbp/s   ({', '.join(f'{t:.0f}' for t in ts)})  "
              f"device memory in use {(tot_b - free_b) / 1e9:.1f} GB", flush=True)


if __name__ == "__main__":
    main()
