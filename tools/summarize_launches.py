#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import sys

src = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else src
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    v = float(r[mv].replace(",", ""))
    v = v / 1e3 if r[mu] == "ns" else v * 1e3 if r[mu] == "ms" else v
    a = agg.setdefault(r[kn][:100], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"# {title}")
print("# ncu --metrics gpu__time_duration.sum --clock-control none ; per-launch times are cold-cache and serialised: compare SHARES")
print(f"# total {tot / 1e3:.2f} ms over {len(data)} launches")
print("launches     total_us    mean_us  share  kernel")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:8d} {t:12.1f} {t / n:10.2f} {t / tot:6.3f}  {name}")
