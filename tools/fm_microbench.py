#!/usr/bin/env python3
"""FM backward-search microbenchmark (BASELINE.json config 5): random k-mer queries, k = 17..31, against a synthetic BWT.

    python tools/fm_microbench.py [--symbols 4000000000] [--strings 500000] [--queries 100000000] [--k0 13] [--cpu-symbols 268435456]

Prints one JSON line per (k, prefix table off/on) with queries/s, executed updateInterval steps, algorithmic bytes
(steps x 2 rank queries x 32 B) and the achieved GB/s, then one line with the reference's own findInterval under
`omp parallel for` (oracle/_ref/fm_dump --time) on a smaller synthetic BWT written in the reference's file format.
Queries and the BWT live in HBM before the timed region; times are CUDA events around the kernel (pbsc_findinterval_device).
"""
import argparse
import ctypes as C
import json
import os
import struct
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from longreadselfcorrect_b200 import api  # noqa: E402


def device_search(idx, which, kmers, k, reps=3):
    n = kmers.numel()
    lo = torch.empty(n, dtype=torch.int64, device="cuda")
    hi = torch.empty(n, dtype=torch.int64, device="cuda")
    st = torch.empty(n, dtype=torch.uint8, device="cuda")
    ms = C.c_float(0)
    best = None
    for _ in range(reps + 1):
        rc = api.lib().pbsc_findinterval_device(idx._h, C.c_int(which), C.c_void_p(kmers.data_ptr()), C.c_int(k), C.c_uint64(n),
                                                C.c_void_p(lo.data_ptr()), C.c_void_p(hi.data_ptr()), C.c_void_p(st.data_ptr()), C.byref(ms))
        api._check(rc)
        best = ms.value if best is None else min(best, ms.value)
    return best, lo, hi, st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--symbols", type=int, default=4_000_000_000)
    ap.add_argument("--strings", type=int, default=500_000)
    ap.add_argument("--queries", type=int, default=100_000_000)
    ap.add_argument("--k0", type=int, default=13)
    ap.add_argument("--cpu-symbols", type=int, default=1 << 28)
    ap.add_argument("--cpu-queries", type=int, default=4_000_000)
    ap.add_argument("--ks", default="17,19,21,23,25,27,29,31", help="k-mer lengths to run")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--json", action="store_true", help="one summary JSON object on the last line (what bench.py embeds as sub_results.fm_microbench); "
                                                        "skips the CPU leg")
    args = ap.parse_args()
    torch.cuda.set_device(args.device)
    summary = {"metric": "FM k-mer backward-search queries/s", "bwt_symbols": args.symbols, "queries_per_k": args.queries, "strings": args.strings,
               "what": "BASELINE.json configs[4]: uniform random k-mers against a synthetic i.i.d. BWT resident in HBM, CUDA events around "
                       "findinterval_packed_kernel; algorithmic bytes = executed updateInterval steps x 2 rank queries x 32 B; issued sectors = "
                       "with the prefix table one 32-byte entry replaces the first k0 - 1 steps", "results": []}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    t = time.time()
    idx = api.Index.synthetic(args.symbols, args.strings, 5, device=args.device)
    print(f"[fm] synthetic BWT {args.symbols} symbols x 2 strands on device in {time.time() - t:.1f}s, {idx.device_bytes() / 1e9:.2f} GB", file=sys.stderr)
    g = torch.Generator(device="cuda")
    g.manual_seed(105)
    for k in [int(x) for x in args.ks.split(",")]:
        kmers = torch.randint(0, 1 << 62, (args.queries,), generator=g, device="cuda", dtype=torch.int64) & ((1 << (2 * k)) - 1)
        ref = None
        for k0 in (0, args.k0):
            idx.build_prefix_table(k0)
            ms, lo, hi, st = device_search(idx, api.PBSC_BWT, kmers, k)
            if k0 == 0:
                steps = int(st.to(torch.int64).sum())
                ref = (lo.clone(), hi.clone())
                valid = lo <= hi
            else:
                # same intervals wherever the k-mer occurs; empty intervals stay empty
                assert bool(torch.equal(lo[valid], ref[0][valid])) and bool(torch.equal(hi[valid], ref[1][valid])) and bool((lo[~valid] > hi[~valid]).all())
            alg_bytes = steps * 64.0
            # sectors the kernel asks for: every executed step reads the sector(s) of its two bounds; with the table the first
            # k0 - 1 steps of every query are one 32-byte entry instead (executed steps beyond them are unchanged)
            issued = float(2 * steps) if k0 == 0 else float(args.queries + 2 * int((st.to(torch.int64) - (k0 - 1)).clamp_(min=0).sum()))
            rec = {"metric": "FM k-mer backward-search queries/s", "k": k, "prefix_k0": k0, "queries": args.queries,
                   "bwt_symbols": args.symbols, "ms": ms, "value": args.queries / (ms / 1e3), "unit": "queries/s",
                   "update_steps": steps, "steps_per_query": steps / args.queries,
                   "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": alg_bytes / (ms / 1e3) / 1e9 / peak, "issued_sectors": issued, "issued_GBs": issued * 32.0 / (ms / 1e3) / 1e9}}
            summary["results"].append({"k": k, "prefix_k0": k0, "Gq_per_s": rec["value"] / 1e9, "ms": ms, "steps_per_query": rec["steps_per_query"],
                                       "algorithmic_GBs": rec["roofline"]["achieved"], "frac_of_streaming_peak": rec["roofline"]["frac"],
                                       "issued_GBs": rec["roofline"]["issued_GBs"]})
            if not args.json:
                print(json.dumps(rec), flush=True)
        del kmers
    idx.close()
    if args.json:
        try:
            summary["random_sector_peak_GBs"] = api.random_sector_peak(args.symbols // 2, args.device)
            summary["note"] = ("issued_GBs is an upper bound of the sector requests (two per executed step, one per prefix-table entry); it is NOT DRAM traffic "
                               "and is not set against random_sector_peak_GBs: the first ~9 steps of a plain search touch at most 2 x 4^t distinct sectors "
                               "(L2-resident), and both bounds of a small interval share a sector")
        except Exception:
            pass
        print(json.dumps(summary), flush=True)
        return
    # ---- reference CPU baseline on a smaller synthetic BWT in the reference's own file format ----
    fm_dump = os.path.join(ROOT, "oracle", "_ref", "fm_dump")
    if os.path.exists(fm_dump) and args.cpu_symbols > 0:
        small = api.Index.synthetic(args.cpu_symbols, max(1, args.strings * args.cpu_symbols // max(args.symbols, 1)), 5)
        n = small.num_symbols(api.PBSC_BWT)
        rank = np.zeros(256, dtype=np.uint8)
        for i, c in enumerate(b"$ACGT"):
            rank[c] = i
        with tempfile.TemporaryDirectory() as d:
            runs = bytearray()
            chunk = 1 << 26
            carry_sym, carry_len = -1, 0
            out = []
            for first in range(0, n, chunk):
                cnt = min(chunk, n - first)
                r = rank[np.frombuffer(small.symbols(api.PBSC_BWT, first, cnt), dtype=np.uint8)]
                change = np.flatnonzero(np.diff(r)) + 1
                starts = np.concatenate(([0], change))
                lens = np.diff(np.concatenate((starts, [cnt])))
                syms = r[starts]
                # merge with the run carried over from the previous chunk
                if carry_sym == int(syms[0]):
                    lens[0] += carry_len
                elif carry_sym >= 0:
                    syms = np.concatenate(([carry_sym], syms)); lens = np.concatenate(([carry_len], lens))
                carry_sym, carry_len = int(syms[-1]), int(lens[-1])
                syms, lens = syms[:-1], lens[:-1]
                reps = (lens + 30) // 31
                rid = np.repeat(np.arange(syms.size), reps)
                firstc = np.cumsum(reps) - reps
                within = np.arange(rid.size) - firstc[rid]
                clen = np.where(within == reps[rid] - 1, lens[rid] - 31 * (reps[rid] - 1), 31)
                out.append(((syms[rid].astype(np.uint16) << 5) | clen.astype(np.uint16)).astype(np.uint8))
            l = carry_len
            tail = []
            while l > 0:
                c = min(l, 31); tail.append((carry_sym << 5) | c); l -= c
            out.append(np.array(tail, dtype=np.uint8))
            runs = np.concatenate(out)
            path = os.path.join(d, "syn.bwt")
            with open(path, "wb") as f:
                f.write(struct.pack("<HQQQi", 0xCACA, small.num_strings(api.PBSC_BWT), n, runs.size, 0))
                f.write(runs.tobytes())
            rng = np.random.Generator(np.random.PCG64(105))
            cores = os.cpu_count() or 1
            for k in (19, 31):
                q = rng.integers(0, 4, size=(args.cpu_queries, k), dtype=np.uint8)
                qf = os.path.join(d, f"q{k}.txt")
                with open(qf, "wb") as f:
                    f.write(b"\n".join(np.frombuffer(b"ACGT", dtype=np.uint8)[q].view(f"S{k}").ravel().tolist()) + b"\n")
                r = subprocess.run([fm_dump, "--time", str(cores), path, qf], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                f = r.stdout.split()
                if len(f) >= 4:
                    print(json.dumps({"impl": "reference", "metric": "FM k-mer backward-search queries/s", "k": k, "queries": int(f[0]),
                                      "bwt_symbols": n, "seconds": float(f[1]), "value": float(f[2]) * 1e6, "unit": "queries/s", "cores": cores,
                                      "update_steps": int(f[3]), "kind": "reference (BWTAlgorithms::findInterval under omp parallel for)"}), flush=True)
        small.close()


if __name__ == "__main__":
    main()
