#!/usr/bin/env python3
"""Per-source-line executed instructions and average active threads from an .ncu-rep (divergence hunting).
    python tools/ncu_lines.py REPORT.ncu-rep [n_lines]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True).stdout
cur, hdr, agg = None, None, {}
for r in csv.reader(io.StringIO(txt)):
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 2 and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if not hdr or len(r) < len(hdr):
        continue
    try:
        ie = float(r[hdr["Instructions Executed"]] or 0)
        te = float(r[hdr["Thread Instructions Executed"]] or 0)
        sm = float(r[hdr["# Samples"]] or 0)
    except ValueError:
        continue
    a = agg.setdefault((cur, r[0]), [0.0, 0.0, 0.0, r[1][:110]])
    a[0] += ie; a[1] += te; a[2] += sm
ti = sum(a[0] for a in agg.values()) or 1
tt = sum(a[1] for a in agg.values())
print(f"warp instructions {ti:.3g}, thread instructions {tt:.3g}, average active threads {tt / ti:.2f}")
print("share of warp instructions | avg active threads | share of samples | line")
ts = sum(a[2] for a in agg.values()) or 1
for (f, l), (ie, te, sm, src) in sorted(agg.items(), key=lambda x: -x[1][0])[:n]:
    print(f"  {100 * ie / ti:5.2f}%  {te / ie if ie else 0:5.1f}  {100 * sm / ts:5.2f}%  {f}:{l}  {src}")
