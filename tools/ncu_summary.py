#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel:
launch count, total / average duration and share of the captured GPU time.

    python tools/ncu_summary.py gpurun_out/launches.csv > profiles/rNN_launches.md
"""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0, "", ""])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}[row["Metric Unit"]]
        a = agg[k]
        a[0] += 1; a[1] += v; a[2] = row["Block Size"]; a[3] = row["Grid Size"]
    tot = sum(v[1] for v in agg.values()) or 1.0
    print(f"source: {path}  (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)")
    print()
    print("| kernel | launches | total ms | avg us | share | block | grid (last) |")
    print("|---|---:|---:|---:|---:|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {v[0]} | {v[1] / 1e6:.3f} | {v[1] / v[0] / 1e3:.1f} | {v[1] / tot:.3f} | {v[2]} | {v[3]} |")


if __name__ == "__main__":
    main(sys.argv[1])
