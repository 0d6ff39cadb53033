#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel (cold-cache, serialised)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
tot = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if "Metric Value" not in row or row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    k = re.sub(r"\(.*", "", row["Kernel Name"])
    tot[k][0] += 1
    tot[k][1] += v
allt = sum(v for _, v in tot.values())
print(f"{'ms':>10} {'share':>6} {'launches':>8} {'avg us':>9}  kernel")
for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{v / 1e3:10.3f} {100 * v / allt:5.1f}% {n:8d} {v / n:9.1f}  {k[:110]}")
print(f"total {allt / 1e3:.2f} ms over {sum(n for n, _ in tot.values())} launches")
