#!/usr/bin/env python
"""Static SASS check before spending GPU time: for one kernel of a .so, list the loops (backward branches) with
their instruction counts and the local-memory (spill) instructions inside them.
Usage: sass_loops.py lib.so <mangled-name-substring>"""
import re
import subprocess
import sys

so, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, funcs = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name:
        continue
    addrs = {a: i for i, (a, _) in enumerate(ins)}
    print(name, len(ins), "instructions,", sum("LDL" in t or "STL" in t for _, t in ins), "local-memory ops")
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addrs:
                loops.append((addrs[tgt], i))
    for lo, hi in sorted(set(loops)):
        body = ins[lo:hi + 1]
        sp = [t for _, t in body if "LDL" in t or "STL" in t]
        mem = sum(1 for _, t in body if re.match(r"(@\S+\s+)?(LDG|STG|LDS|STS|LD\.|ST\.)", t))
        print(f"  loop {ins[lo][0]:#x}-{ins[hi][0]:#x}: {hi - lo + 1:4d} instr, {mem:3d} mem ops, spills: {len(sp)} {sp[:6]}")
