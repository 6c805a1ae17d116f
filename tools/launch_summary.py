#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into the table kept under profiles/.

usage: python tools/launch_summary.py launches.csv "header line" ... > profiles/rNN_launches_X.txt
"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    n = r[ki].split("(")[0]
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) / 1e3  # ns -> us
tot = sum(v[1] for v in agg.values())
for line in sys.argv[2:]:
    print(f"# {line}")
print(f"# total device time over {sum(v[0] for v in agg.values())} launches: {tot / 1e3:.2f} ms\n")
print(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:70]:70s} {c:8d} {t:12.1f} {t / c:10.1f} {100 * t / tot:6.2f}%")
