#!/usr/bin/env python3
"""Digest of an ncu report exported as raw + source CSV (dev tool): key counters, stall mix, hottest instructions."""
import csv
import sys

raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct']
st = [h for h in hdr if 'stall' in h.lower() and 'ratio' in h and 'not_issued' not in h]
for r in rows[2:]:
    print(r[idx['Kernel Name']][:40])
    print('  ' + '  '.join(f"{w.split('.')[0][-28:]}={r[idx[w]]}" for w in want if w in idx))
    vals = []
    for h in st:
        try:
            vals.append((float(r[idx[h]]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
        except Exception:
            pass
    print('   stalls:', ', '.join(f"{h}={v:.2f}" for v, h in sorted(vals, reverse=True)[:7]))
rows = list(csv.reader(open(src)))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
start = hdr_i[0]; end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
hdr = rows[start]; idx = {h: i for i, h in enumerate(hdr)}
key = 'Warp Stall Sampling (All Samples)'


def f(x):
    try:
        return float(x)
    except Exception:
        return 0.0


data = [r for r in rows[start + 1:end] if len(r) > idx[key]]
tot = sum(f(r[idx[key]]) for r in data)
for i, r in sorted(enumerate(data), key=lambda ir: -f(ir[1][idx[key]]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 14]:
    print(f"{f(r[idx[key]]) / tot * 100:6.2f}%  #{i:5d} exec={r[idx['Instructions Executed']]:>9s} thr={r[idx['Avg. Threads Executed']]:>5s}  {r[idx['Source']][:90]}")
mx = max(f(r[idx['Instructions Executed']]) for r in data)
inloop = [f(r[idx[key]]) for r in data if f(r[idx['Instructions Executed']]) > 0.9 * mx]
print("hottest-loop share of samples:", round(sum(inloop) / tot, 3), "n instr", len(inloop))
low = [(f(r[idx[key]]), f(r[idx['Avg. Threads Executed']])) for r in data if f(r[idx['Instructions Executed']]) <= 0.9 * mx]
print("outside: samples share", round(sum(a for a, b in low) / tot, 3), " sample-weighted avg threads", round(sum(a * b for a, b in low) / max(1, sum(a for a, b in low)), 1))
