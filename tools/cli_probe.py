#!/usr/bin/env python3
"""Write a workload's reads + index files to /dev/shm and run the pbcorrect binary on them with PBSC_TRACE=1 (host wall clock
of upload / run / fetch per batch), twice (cold: PREFIX.bwt, --write-fmg; warm: PREFIX.fmg).
    python tools/cli_probe.py [workload] [gpus] [batch_mbp] [lanes]"""
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bench  # noqa: E402
from longreadselfcorrect_b200 import bwt_build  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
gpus = sys.argv[2] if len(sys.argv) > 2 else "1"
batch = sys.argv[3] if len(sys.argv) > 3 else "160"
lanes = sys.argv[4] if len(sys.argv) > 4 else "2"
codes, off = bench.make_data(wl)
n = off.size - 1
d = "/dev/shm/pbsc_cli_probe"
shutil.rmtree(d, ignore_errors=True)
os.makedirs(d)
runs = {}
for ext, rev in (("bwt", False), ("rbwt", True)):
    b = bwt_build.bwt_symbols(codes, off, reverse=rev, device="cuda:0")
    runs[ext] = (bwt_build.run_length_bytes(b), int(b.numel()), n)
    del b
import torch  # noqa: E402
torch.cuda.empty_cache()
bench.write_index_files(os.path.join(d, "idx"), runs, n)
letters, off64 = bench.packed_ascii(codes, off)
bench.write_fasta_ids(os.path.join(d, "reads.fa"), letters, off64, range(n))
for attempt, extra in ((("cold", ["--write-fmg"]), ("warm", [])) if not os.environ.get("CLI_ONE") else (("warm", []),)):
    t0 = time.time()
    r = subprocess.run([bench.PBCORRECT, "pbcorrect", "-t", "8", "--gpus", gpus, "-p", os.path.join(d, "idx"), "-o", os.path.join(d, "out"), "-c", str(wl["c"]), "-g", str(wl["g"]),
                        "--batch-mbp", batch, "--lanes", lanes] + extra + [os.path.join(d, "reads.fa")], env=dict(os.environ, PBSC_TRACE="1", **({"PBSC_ROUND_TRACE": "1"} if os.environ.get("CLI_ROUND_TRACE") else {})),
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    print(f"==== {attempt}: rc {r.returncode}, wall {time.time() - t0:.2f}s ====")
    print(r.stderr[-12000:])
    print(r.stdout[-1200:])
shutil.rmtree(d, ignore_errors=True)
