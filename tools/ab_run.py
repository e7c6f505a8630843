#!/usr/bin/env python3
"""A/B timing of library environment knobs in ONE process (one data set, one BWT build): saves GPU minutes.

    python tools/ab_run.py [--workload cfg2] [--reps 3] [--nodp] [--k0 13] VARIANT [VARIANT ...]

VARIANT is `name` (no variables: the default build) or `name:VAR1=VAL1,VAR2=VAL2`.  For every variant the variables are set,
the index is re-created from the same run-length BWTs (some knobs are read when the index is created), one batch is
uploaded, `reps` passes are timed (CUDA events inside the library) and the median is printed with the phase times of
`pbsc_last_timing`.  Example:

    python tools/ab_run.py base l2_60:PBSC_L2_PERSIST=60 l2_100:PBSC_L2_PERSIST=100 nodpthread:PBSC_DP_THREAD=0 k15:K0=15
"""
import argparse
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse_variant(text):
    name, _, rest = text.partition(":")
    env = {}
    for item in filter(None, rest.split(",")):
        k, eq, v = item.partition("=")
        if not eq:
            raise SystemExit(f"variant {text!r}: expected VAR=VALUE, got {item!r}")
        env[k] = v
    return name, env


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--k0", type=int, default=13)
    ap.add_argument("--nodp", action="store_true")
    ap.add_argument("variants", nargs="+")
    args = ap.parse_args()
    variants = [parse_variant(v) for v in args.variants]

    import bench  # noqa: E402  (repo root)
    from longreadselfcorrect_b200 import api, bwt_build  # noqa: E402

    wl = bench.WORKLOADS[args.workload]
    codes, off = bench.make_data(wl)
    n = off.size - 1
    runs = {}
    for ext, rev in (("bwt", False), ("rbwt", True)):
        b = bwt_build.bwt_symbols(codes, off, reverse=rev, device="cuda:0")
        runs[ext] = (bwt_build.run_length_bytes(b), int(b.numel()), n)
    packed = bench.packed_ascii(codes, off)
    mbp = codes.size / 1e6
    keys = ("seed_ms", "extend_ms", "walk_ms", "dp_ms", "kernel_launches", "dp_thread_rows")
    print(f"workload {args.workload}: {n} reads, {mbp:.1f} Mbp, reps {args.reps}, dp {'off' if args.nodp else 'on'}")
    print(f"{'variant':24s} {'ms':>9s} {'Mbp/s':>8s}  " + " ".join(f"{k:>14s}" for k in keys))
    for name, env in variants:
        env = dict(env)
        k0 = int(env.pop("K0", args.k0))   # `K0=15` in a variant: that variant's short-prefix table length
        saved = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            idx = api.Index.from_runs(runs["bwt"][0], runs["bwt"][1], n, runs["rbwt"][0], runs["rbwt"][1], n)
            if k0:
                idx.build_prefix_table(k0)
            p = api.Params.make(coverage=wl["c"], genome=wl["g"], no_dp=args.nodp)
            batch = api.Batch(idx, p, packed=packed)
            batch.run()   # warm-up: grow-only arenas, learned capacities
            times, last = [], None
            for _ in range(args.reps):
                times.append(batch.run())
                last = api.last_timing()
            ms = statistics.median(times)
            print(f"{name:24s} {ms:9.1f} {mbp / (ms / 1e3):8.1f}  " + " ".join(f"{last[k]:14.1f}" for k in keys), flush=True)
            batch.close()
            idx.close()
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v


if __name__ == "__main__":
    main()
