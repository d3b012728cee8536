#!/usr/bin/env python3
"""Instruction mix of the loops of one kernel from `cuobjdump -sass` output (dev tool)."""
import collections
import re
import sys

path, fn = sys.argv[1], sys.argv[2]
txt = open(path).read().split("Function : ")
body = [t for t in txt if t.split("\n")[0].find(fn) >= 0][0]
ins = []
for l in body.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print(len(ins), "instructions in", fn)
loops = []
for a, t in ins:
    if "BRA" in t:
        mm = re.search(r"0x([0-9a-f]+)", t)
        if mm and int(mm.group(1), 16) < a:
            loops.append((int(mm.group(1), 16), a))
for s, e in sorted(loops, key=lambda x: -(x[1] - x[0]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 3]:
    b = [t for a, t in ins if s <= a <= e]
    hist = collections.Counter()
    for t in b:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        hist[t.split()[0].split(".")[0]] += 1
    fp64 = sum(v for k, v in hist.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print(f"loop {hex(s)}..{hex(e)}: {len(b)} instr, FP64-pipe {fp64}:", sorted(hist.items(), key=lambda kv: -kv[1]))
