#!/usr/bin/env python3
"""Summarise an .ncu-rep: headline metrics, stall-reason mix, opcode mix and the hottest source lines.
    python tools/ncu_summary.py REPORT.ncu-rep [n_lines]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 30


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, vals = raw[0], raw[2] if len(raw) > 2 else raw[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sectors_srcunit_tex_op_read.sum", "dram__sectors_read.sum",
        "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", "sm__icc_request_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h} = {vals[i]} {raw[1][i] if len(raw) > 2 else ''}")

rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "cuda,sass"))))
cur, h2, agg, stall_tot, ops = None, None, {}, {}, {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 2 and r[0] == "Line No":
        h2 = r
        idx = {h: i for i, h in enumerate(h2)}
        stalls = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if not h2 or len(r) < len(h2):
        continue
    if r[0] != "":
        try:
            agg[(cur, int(r[0]))] = (int(r[idx["# Samples"]]), int(r[idx["Instructions Executed"]]), r[1][:100])
        except ValueError:
            pass
    else:
        for s in stalls:
            try:
                stall_tot[s] = stall_tot.get(s, 0) + int(r[idx[s]])
            except ValueError:
                pass
        try:
            t = r[3].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] = ops.get(op, 0) + int(r[idx["Instructions Executed"]])
        except (ValueError, IndexError):
            pass
ts = sum(v[0] for v in agg.values()) or 1
ti = sum(v[1] for v in agg.values()) or 1
print("\nstall reasons (share of samples):")
T = sum(stall_tot.values()) or 1
for s, v in sorted(stall_tot.items(), key=lambda x: -x[1])[:8]:
    print(f"  {s:26s} {100 * v / T:5.1f}%")
print("\nopcode mix:")
TI = sum(ops.values()) or 1
for o, v in sorted(ops.items(), key=lambda x: -x[1])[:12]:
    print(f"  {o:10s} {100 * v / TI:5.1f}%")
byfile = {}
for (f, l), (s, i, _) in agg.items():
    a = byfile.setdefault(f, [0, 0])
    a[0] += s
    a[1] += i
print("\nby file:")
for f, (s, i) in sorted(byfile.items(), key=lambda x: -x[1][0]):
    print(f"  {f:28s} samples {100 * s / ts:5.1f}%  inst {100 * i / ti:5.1f}%")
print("\nhottest lines:")
for (f, l), (s, i, src) in sorted(agg.items(), key=lambda x: -x[1][0])[:nlines]:
    print(f"  {f}:{l:4d} s={100 * s / ts:5.2f}% i={100 * i / ti:5.2f}% {src}")
