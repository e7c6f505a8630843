#!/usr/bin/env python3
"""Index construction on the GPU against the torch stand-in (and, with --reference-mbp, the reference's own `stride index` on
a sample): seconds for both strands of a bench workload, and whether the run-length bytes agree.

    python tools/index_bench.py [--workload cfg2] [--skip-torch] [--reference-mbp 5] [--json]"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--skip-torch", action="store_true")
    ap.add_argument("--reference-mbp", type=float, default=0.0, help="also time oracle/_ref/stride index on the first reads up to this many Mbp")
    ap.add_argument("--json", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    from longreadselfcorrect_b200 import api, bwt_build

    wl = bench.WORKLOADS[args.workload]
    codes, off = bench.make_data(wl)
    packed = bench.packed_ascii(codes, off)
    mbp = codes.size / 1e6
    res = {"workload": args.workload, "mbp": mbp, "reads": int(off.size - 1)}
    api.build_bwt((packed[0][:int(off[8])], packed[1][:9]))   # context, module load
    t0 = time.perf_counter()
    mine = [api.build_bwt(packed, reverse=rev) for rev in (False, True)]
    res["gpu_builder_s"] = time.perf_counter() - t0
    res["gpu_builder_mbp_per_s"] = mbp / res["gpu_builder_s"]
    res["run_length_bytes"] = [int(m[0].size) for m in mine]
    if not args.skip_torch:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        theirs = [bwt_build.run_length_bytes(bwt_build.bwt_symbols(codes, off, reverse=rev, device="cuda:0")) for rev in (False, True)]
        torch.cuda.synchronize()
        res["torch_builder_s"] = time.perf_counter() - t0
        res["identical_to_torch_builder"] = all(np.array_equal(m[0], np.asarray(t)) for m, t in zip(mine, theirs))
    if args.reference_mbp > 0 and os.path.exists(bench.REF_STRIDE):
        n = int(np.searchsorted(off, args.reference_mbp * 1e6, side="right")) - 1
        n = max(n, 1)
        letters = np.frombuffer(b"ACGT", dtype=np.uint8)[codes[:int(off[n])]]
        with tempfile.TemporaryDirectory() as d:
            with open(os.path.join(d, "s.fa"), "wb") as f:
                for i in range(n):
                    f.write(b">r%d\n" % i + letters[int(off[i]):int(off[i + 1])].tobytes() + b"\n")
            t0 = time.perf_counter()
            subprocess.run([bench.REF_STRIDE, "index", "-t", str(os.cpu_count() or 1), "s.fa"], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            secs = time.perf_counter() - t0
            sub = (packed[0][:int(off[n])], packed[1][:n + 1])
            t0 = time.perf_counter()
            api.build_index_files(sub, os.path.join(d, "g"))
            gsecs = time.perf_counter() - t0
            same = all(open(os.path.join(d, f"s.{e}"), "rb").read() == open(os.path.join(d, f"g.{e}"), "rb").read() for e in ("bwt", "rbwt", "sai", "rsai"))
        smbp = float(off[n]) / 1e6
        res["reference_sample"] = {"mbp": smbp, "reads": n, "stride_index_s": secs, "stride_index_mbp_per_s": smbp / secs, "threads": os.cpu_count(),
                                   "gpu_builder_s": gsecs, "files_identical": same}
    print(json.dumps(res) if args.json else "\n".join(f"{k}: {v}" for k, v in res.items()), flush=True)


if __name__ == "__main__":
    main()
