#!/usr/bin/env python3
"""Region-level view of an ncu source page (dev tool): tools/ncu_src_profile.py <src.csv> <kernel index> [min share]
Consecutive SASS instructions with (nearly) the same execution count form a segment; per segment: share of the warp-stall
samples, instructions, executions per instruction, FP64 share, average active threads, leading stall reasons."""
import csv
import sys

path, kidx = sys.argv[1], int(sys.argv[2])
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
rows = list(csv.reader(open(path)))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
start = hdr_i[kidx]
end = hdr_i[kidx + 1] - 1 if kidx + 1 < len(hdr_i) else len(rows)
hdr = rows[start]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]


def f(x):
    try:
        return float(x)
    except Exception:
        return 0.0


data = [r for r in rows[start + 1:end] if len(r) > ix['Instructions Executed']]
tot = sum(f(r[ix['Warp Stall Sampling (All Samples)']]) for r in data) or 1.0
tot_exec = sum(f(r[ix['Instructions Executed']]) for r in data)
print(f"{len(data)} SASS instructions, {tot:.0f} samples, {tot_exec:.4g} warp instructions executed")
segs, cur = [], []
for i, r in enumerate(data):
    e = f(r[ix['Instructions Executed']])
    if cur:
        e0 = f(cur[0][1][ix['Instructions Executed']])
        if abs(e - e0) > 0.15 * max(e, e0, 1.0):
            segs.append(cur)
            cur = []
    cur.append((i, r))
if cur:
    segs.append(cur)
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
for sg in segs:
    smp = sum(f(r[ix['Warp Stall Sampling (All Samples)']]) for _, r in sg)
    ex = sum(f(r[ix['Instructions Executed']]) for _, r in sg)
    if smp / tot < min_share and ex / tot_exec < min_share:
        continue
    n = len(sg)
    n64 = sum(1 for _, r in sg if any(op in r[ix['Source']] for op in FP64))
    thr = sum(f(r[ix['Avg. Threads Executed']]) * f(r[ix['Instructions Executed']]) for _, r in sg) / max(ex, 1.0)
    st = sorted(((sum(f(r[ix[c]]) for _, r in sg), c.replace('stall_', '')) for c in stall_cols), reverse=True)[:4]
    sst = sum(v for v, _ in sorted(((sum(f(r[ix[c]]) for _, r in sg), c) for c in stall_cols), reverse=True)) or 1.0
    print(f"#{sg[0][0]:5d}-{sg[-1][0]:5d} n={n:4d} fp64={n64:4d} exec/instr={ex / n:11.0f} exec share={ex / tot_exec:6.3f} samples={smp / tot:6.3f} thr={thr:5.1f}  "
          + ", ".join(f"{c}={v / sst:.2f}" for v, c in st) + "   | " + sg[0][1][ix['Source']][:50])
