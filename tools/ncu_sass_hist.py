#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source sass` output: opcode mix (weighted by
executed instructions) and the hottest stall sites.  Usage: ncu_sass_hist.py sass.csv [kernel_idx]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
# split into kernels: a row starting with "Kernel Name" begins one
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
k = kernels[which]
hdr = k["hdr"]
ia, isrc, ie, isamp = (hdr.index(x) for x in ("Address", "Source", "Instructions Executed", "# Samples"))
tot = sum(int(r[ie]) for r in k["rows"])
ts = sum(int(r[isamp]) for r in k["rows"])
print(k["name"], "| sass lines", len(k["rows"]), "| inst executed", tot, "| samples", ts)
c, cs = Counter(), Counter()
for r in k["rows"]:
    toks = r[isrc].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    c[op] += int(r[ie])
    cs[op] += int(r[isamp])
for op, n in c.most_common(28):
    print(f"  {op:10s} {n / tot * 100:5.1f}% inst   {cs[op] / ts * 100:5.1f}% samples")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = Counter()
for r in k["rows"]:
    for h in stalls:
        v = r[hdr.index(h)]
        if v:
            agg[h] += int(v)
print("stall totals:", ", ".join(f"{h[6:]}={v / max(1, sum(agg.values())) * 100:.0f}%" for h, v in agg.most_common(8)))
print("hottest instructions by samples:")
for r in sorted(k["rows"], key=lambda r: -int(r[isamp]))[:25]:
    top = sorted(((int(r[hdr.index(h)] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"  {int(r[isamp]) / ts * 100:5.1f}%  {r[isrc][:70]:70s} {top}")
