#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, launches,
total and mean duration (us).  Usage: launch_list.py file.csv [skip_first_n]"""
import csv
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    for r in csv.reader(f):
        if len(r) > 14 and r[12] == "gpu__time_duration.sum":
            rows.append((r[4].split("(")[0], r[8], r[7], float(r[14].replace(",", ""))))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = OrderedDict()
for name, grid, block, ns in rows:
    a = agg.setdefault((name, grid, block), [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':58s} {'grid':>14s} {'block':>12s} {'n':>5s} {'total_us':>11s} {'mean_us':>10s} {'share':>6s}")
for (name, grid, block), (n, ns) in agg.items():
    print(f"{name[:58]:58s} {grid:>14s} {block:>12s} {n:5d} {ns / 1e3:11.1f} {ns / 1e3 / n:10.1f} {ns / tot:6.1%}")
