#!/usr/bin/env python
"""Per-CUDA-source-line instruction counts / stall samples from
`ncu --page source --csv --print-source cuda,sass`.  Usage: ncu_src_hist.py cs.csv [topN]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, agg, first_fn = None, None, {}, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        if first_fn is None:
            first_fn = r[1]
        fn = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-" and fn == first_fn:   # a CUDA source line row
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, [0, 0, r[1]])
        a[0] += int(r[ie] or 0)
        a[1] += int(r[isamp] or 0)
tot = sum(a[0] for a in agg.values())
ts = sum(a[1] for a in agg.values())
print(first_fn, "inst", tot, "samples", ts)
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0] / tot * 100:5.1f}% inst {a[1] / max(ts,1) * 100:5.1f}% smp  {f}:{ln:<4d} {a[2].strip()[:95]}")
