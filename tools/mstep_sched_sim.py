#!/usr/bin/env python3
"""Host-side simulation of the M-step's lane scheduler (profiles/r02_ncu_summary.md §2): 2 368 warps of 32 lanes
(148 SMs × 4 blocks × 4 warps) solve 10⁶ EM instances whose chain lengths follow the bench workload; a lane keeps an instance for
1–6 derivative sweeps, a warp's sweep costs the longest chain among its busy lanes (+ OV groups of bookkeeping).  Policies:
  global   one queue sorted by chain length, longest first (what k_mstep_ws does)
  batch B  a warp reserves 32·B consecutive instances at a time and its lanes steal inside the batch
  chunks   per-block queues of equal work (first 1−eps of the list) + one shared overflow queue
  chunks+neighbour   per-block queues of equal work, lanes of a drained block steal from the next block's queue
Prints (lane efficiency inside the sweeps, useful / capacity over the whole launch, launch length in group-steps).
Usage: python tools/mstep_sched_sim.py      (CPU only, a few minutes)"""
import heapq
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpelnano_b200 as cpn  # noqa: E402

rng = np.random.default_rng(0)
nb = cpn.simulations.make_nano_batch(20000, M=2, seed=1)
Lr = np.diff(nb.grp_off)
print("chain lengths: mean %.1f sd %.1f min %d max %d" % (Lr.mean(), Lr.std(), Lr.min(), Lr.max()))
R, S = 100000, 10
Linst = np.repeat(np.sort(rng.choice(Lr, R))[::-1], S)
N = len(Linst)
ns = rng.choice([1, 2, 3, 4, 5, 6], N, p=[.15, .35, .25, .12, .08, .05])        # sweeps per solve (mean 2.8)
W, OV = 2368, 6.0


def run(take, Wn=W):
    t = [(0.0, w) for w in range(Wn)]
    heapq.heapify(t)
    lanesL = np.zeros((Wn, 32), int); lanesN = np.zeros((Wn, 32), int)
    useful = busy = tend = 0.0
    while t:
        tm, w = heapq.heappop(t)
        for l in range(32):
            if lanesN[w, l] == 0:
                i = take(w)
                if i is not None:
                    lanesL[w, l] = Linst[i]; lanesN[w, l] = ns[i]
        act = lanesN[w] > 0
        if not act.any():
            tend = max(tend, tm)
            continue
        mx = lanesL[w][act].max()
        useful += lanesL[w][act].sum(); busy += mx
        lanesN[w][act] -= 1
        heapq.heappush(t, (tm + mx + OV, w))
    return round(useful / (32 * busy), 3), round(useful / (32 * Wn * tend), 3), tend


def policy_global():
    head = [0]

    def take(w):
        if head[0] < N:
            head[0] += 1
            return head[0] - 1
        return None
    return take


def policy_batch(B):
    head = [0]
    loc = [[0, 0] for _ in range(W)]

    def take(w):
        if loc[w][0] >= loc[w][1] and head[0] < N:
            loc[w] = [head[0], min(N, head[0] + 32 * B)]
            head[0] = loc[w][1]
        if loc[w][0] < loc[w][1]:
            loc[w][0] += 1
            return loc[w][0] - 1
        return None
    return take


def policy_chunks(eps, NB=592, WPB=4, neighbour=False):
    work = np.cumsum(Linst + OV)
    tot = work[-1] * (1 - (0 if neighbour else eps))
    bounds = np.concatenate([[0], np.searchsorted(work, np.linspace(0, tot, NB + 1)[1:], side="left")])
    if neighbour:
        bounds[-1] = N
    chead, cend = bounds[:-1].copy(), bounds[1:].copy()
    ghead = [int(bounds[-1])]
    cur = np.arange(NB * WPB) // WPB

    def take(w):
        b = cur[w]
        k = 0
        while chead[b] >= cend[b] and neighbour and k < NB:
            b = (b + 1) % NB; k += 1
        if chead[b] < cend[b]:
            cur[w] = b
            chead[b] += 1
            return chead[b] - 1
        if not neighbour and ghead[0] < N:
            ghead[0] += 1
            return ghead[0] - 1
        return None
    return take


print("global", run(policy_global()))
for B in (2, 4, 8, 16):
    print("batch", B, run(policy_batch(B)))
for eps in (0.05, 0.1, 0.15, 0.25):
    print("chunks eps", eps, run(policy_chunks(eps)))
print("chunks + neighbour stealing", run(policy_chunks(0, neighbour=True)))
