#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` output: per-kernel stall totals + hottest SASS lines.
usage: ncu -i rep --page source --csv [--kernel-name regex:..] | python tools/ncu_stalls.py [ntop]"""
import csv, sys
ntop = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
# the export holds one block per kernel: a "Kernel Name" line, a header line, then instruction lines
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}; blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["data"].append(r)
for b in blocks[:1] if "--all" not in sys.argv else blocks:
    hdr, data = b["hdr"], b["data"]
    S, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[S]) for r in data)
    print("kernel:", b["name"][:110]); print("total samples", tot, " SASS lines", len(data))
    agg = {}
    for r in data:
        for i in stalls: agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
    print("stall totals:", [(k, v, f"{100*v/max(tot,1):.0f}%") for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]])
    top = sorted(range(len(data)), key=lambda i: -int(data[i][S]))[:ntop]
    for i in sorted(top):
        r = data[i]
        st = {hdr[j]: int(r[j]) for j in stalls if int(r[j]) > 0}
        print(f"{i:5d} {r[src].strip()[:64]:64s} smp={r[S]:>6s} exec={r[ie]:>8s}", sorted(st.items(), key=lambda x: -x[1])[:3])
