#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/launch_summary.py launches.csv [top]"""
import collections
import csv
import sys

UNIT = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}


def summarize(path, top=20):
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in csv.reader(open(path)):
        if len(r) > 3 and r[0] == "ID":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0].replace("pbsc::", "")
        agg[name][0] += 1
        agg[name][1] += float(d["Metric Value"].replace(",", "")) * UNIT[d["Metric Unit"]]
    total = sum(t for _, t in agg.values())
    lines = [f"total {total:.2f} ms over {sum(n for n, _ in agg.values())} launches"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        lines.append(f"{t:10.2f} ms {100 * t / total:5.1f}% {n:5d}  {k}")
    return "\n".join(lines)


if __name__ == "__main__":
    print(summarize(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 20))
