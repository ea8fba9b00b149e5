#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, share, mean."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
agg = collections.defaultdict(lambda: [0, 0.0, ""])
for r in rows[1:]:
    if r[vi] == "": continue
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    k = r[ki].replace("void unnamed>::", "").split("(")[0]
    agg[k][0] += 1; agg[k][1] += v; agg[k][2] = r[gi]
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':48s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'mean us':>9s}  grid")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:48s} {v[0]:8d} {v[1]/1e3:10.2f} {100*v[1]/tot:6.1f}% {v[1]/v[0]:9.1f}  {v[2]}")
