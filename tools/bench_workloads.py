#!/usr/bin/env python3
"""Throughput of the two comparison workloads of BASELINE.json (configs 4 and 5) on one GPU, through the C ABI.

    python tools/bench_workloads.py --workload group      [--regions 20000]
    python tools/bench_workloads.py --workload two_sample [--regions 8] [--perms 1000]

group (grp_perm_tests.jl:200-291, :432-499): 6 vs 6 fitted models per region → summaries of the 12 models
(cpn_stat_sums_batch), CMD of all 66 pairs (cpn_cmd_batch), all 924 labelings (cpn_grp_unmat_sweep) and the 64 sign
patterns of the matched test (cpn_grp_mat_sweep).  two_sample (two_samp_perm_tests.jl:41-87, :157-273): per region
P permutations of 30+30 reads → 2P multi-start EM fits, 2P summaries, P CMDs, p-values by counting.

Not a bench.py line (the contract's metric is config 2); the JSON printed here is kept under profiles/.
"""
import argparse
import json
import os
import sys
import time
from itertools import combinations

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cpelnano_b200 as cpn  # noqa: E402


def sub_regions(lib, nb, size=350):
    subs = [lib.nls_reg_info(int(nb.chrst[r]), int(nb.chrend[r]), size, nb.cpg_pos[nb.cpg_off[r]:nb.cpg_off[r + 1]]) for r in range(nb.R)]
    sub_off = np.concatenate([[0], np.cumsum([len(s[0]) for s in subs])]).astype(np.int64)
    return sub_off, np.concatenate([s[2] for s in subs]), np.concatenate([s[3] for s in subs])


def group(args):
    eng = cpn.Engine()
    lib = eng.lib
    R, n1, n2 = args.regions, 6, 6
    S = n1 + n2
    nb = cpn.simulations.make_nano_batch(R, M=1, seed=5)          # layouts only; the reads are not used
    sub_off, sub_p, sub_q = sub_regions(lib, nb)
    rng = np.random.default_rng(6)
    # gen_grp_comp_data (simulations.jl:588-660): ϕ_s = ϕ_group + nse·[U(−.1,.1), U(−.5,.5), U(0,.2)]; 10 % of regions differ
    base = np.stack([rng.uniform(-1, 1, R), rng.uniform(-20, 20, R), rng.uniform(0, 5, R)], axis=1)
    shift = np.where(rng.random(R)[:, None] < 0.1, np.array([[0.5, 10.0, 1.0]]), 0.0)
    phi = np.empty((R, S, 3))
    for s in range(S):
        noise = np.stack([rng.uniform(-.1, .1, R), rng.uniform(-.5, .5, R), rng.uniform(0, .2, R)], axis=1)
        phi[:, s] = base + (shift if s >= n1 else 0.0) + noise
    reg_of = np.repeat(np.arange(R, dtype=np.int64), S)
    stages = {}

    def timed(name, f):
        t0 = time.perf_counter()
        out = f()
        stages[name] = {"wall_ms": (time.perf_counter() - t0) * 1e3, "device_ms": lib.last_device_ms()}
        return out

    ss = timed("stat_sums_12_models", lambda: lib.stat_sums_batch(phi.reshape(-1, 3), nb.cpg_off, nb.rho_n, nb.d_n, sub_off, sub_p, sub_q,
                                                                  reg_of=reg_of))
    ii, jj = np.array([(i, j) for i in range(S) for j in range(i + 1, S)]).T
    pa = (np.arange(R)[:, None] * S + ii[None]).reshape(-1)
    pb = (np.arange(R)[:, None] * S + jj[None]).reshape(-1)
    cmd, cmd_off = timed("cmd_66_pairs", lambda: lib.cmd_batch(pa, pb, phi.reshape(-1, 3), reg_of, ss["ex_off"], ss["ex"], ss["exx"], ss["k_off"],
                                                                ss["nme"], nb.cpg_off, nb.rho_n, nb.d_n, sub_off, sub_p, sub_q))
    K = np.diff(sub_off)
    t0 = time.perf_counter()
    tbl_off = sub_off
    # cmd table [R][S][S][K_r], NaN outside the upper triangle (layout of cpn_grp_unmat_sweep)
    tbl = np.full(int((K * S * S).sum()), np.nan)
    base_r = np.concatenate([[0], np.cumsum(K * S * S)])[:-1]
    npair = len(ii)
    for k in range(int(K.max())):
        sel = np.nonzero(K > k)[0]
        src = cmd_off[(sel[:, None] * npair + np.arange(npair)[None])] + k
        dst = base_r[sel][:, None] + (ii[None] * S + jj[None]) * K[sel][:, None] + k
        tbl[dst.reshape(-1)] = cmd[src.reshape(-1)]
    mml, nme = ss["mml"], ss["nme"]                             # already [R][S][K_r]
    stages["host_table"] = {"wall_ms": (time.perf_counter() - t0) * 1e3, "device_ms": 0.0}
    labelings = np.array(list(combinations(range(1, S + 1), n1)), dtype=np.int32)
    T, cnt, p = timed("unmat_sweep_924", lambda: lib.grp_unmat_sweep(n1, n2, tbl_off, mml, nme, tbl, labelings, True))
    # matched: models s and n1+s are a pair; split the [R][S][K_r] tables by group
    t0 = time.perf_counter()
    s_of_entry = np.repeat(np.arange(R * S) % S, np.diff(ss["k_off"]))
    g1, g2 = s_of_entry < n1, s_of_entry >= n1
    g1m, g2m, g1n, g2n = ss["mml"][g1], ss["mml"][g2], ss["nme"][g1], ss["nme"][g2]
    stages["host_split"] = {"wall_ms": (time.perf_counter() - t0) * 1e3, "device_ms": 0.0}
    Tm, cm, pm = timed("mat_sweep_64", lambda: lib.grp_mat_sweep(n1, tbl_off, g1m, g2m, g1n, g2n, np.arange(64, dtype=np.int64), True))
    dev = sum(v["device_ms"] for v in stages.values())
    wall = sum(v["wall_ms"] for v in stages.values())
    sig = float(np.nanmean(p[: 3 * int(sub_off[-1])] < 0.05))
    # the same through the fused entry point (ϕ̂ in, statistics out)
    fused = {}
    for rep in range(2):                                                     # first pass warms the pool up
        lib.set_profile(True)
        t0 = time.perf_counter()
        Tf, cf, pf = lib.grp_test_batch(n1, n2, False, phi, nb.cpg_off, nb.rho_n, nb.d_n, sub_off, sub_p, sub_q, labelings, True)
        fused["unmatched"] = {"wall_ms": (time.perf_counter() - t0) * 1e3, "device_ms": lib.last_device_ms(),
                              "kernel_ms": {k: round(v[0], 3) for k, v in lib.get_profile().items() if v[1]}}
        lib.set_profile(True)
        t0 = time.perf_counter()
        Tg, cg, pg, tcg = lib.grp_test_batch(n1, n2, True, phi, nb.cpg_off, nb.rho_n, nb.d_n, sub_off, sub_p, sub_q, np.arange(64), True)
        fused["matched"] = {"wall_ms": (time.perf_counter() - t0) * 1e3, "device_ms": lib.last_device_ms(),
                            "kernel_ms": {k: round(v[0], 3) for k, v in lib.get_profile().items() if v[1]}}
        lib.set_profile(False)
    same = bool(np.array_equal(Tf, T, equal_nan=True) and np.array_equal(cf, cnt) and np.array_equal(Tg, Tm, equal_nan=True) and np.array_equal(cg, cm))
    fw = fused["unmatched"]["wall_ms"] + fused["matched"]["wall_ms"]
    fd = fused["unmatched"]["device_ms"] + fused["matched"]["device_ms"]
    print(json.dumps({"workload": "group 6 vs 6, unmatched (924 labelings) + matched (64 sign patterns)", "regions": R, "models": R * S,
                      "pairs": int(len(pa)), "sub_regions": int(sub_off[-1]),
                      "staged": {"stages": stages, "device_ms": dev, "wall_ms": wall, "regions_per_s_device": R / (dev * 1e-3),
                                 "regions_per_s_wall": R / (wall * 1e-3),
                                 "note": "one C-ABI call per stage; wall includes numpy packing and pageable host↔device copies"},
                      "fused": {"calls": fused, "device_ms": fd, "wall_ms": fw, "regions_per_s_device": R / (fd * 1e-3),
                                "regions_per_s_wall": R / (fw * 1e-3), "identical_to_staged": same,
                                "unmatched_only_regions_per_s_wall": R / (fused["unmatched"]["wall_ms"] * 1e-3),
                                "note": "cpn_grp_test_batch: phi in, (T, cnt, p) out; device_ms includes the copies"},
                      "frac_p_lt_0.05": sig}))


def two_sample(args):
    eng = cpn.Engine()
    lib = eng.lib
    R, P, n1 = args.regions, args.perms, 30
    cfg = cpn.CpelNanoConfig()
    nb = cpn.simulations.make_nano_batch(R, M=2 * n1, seed=9)
    sub_off, sub_p, sub_q = sub_regions(lib, nb)
    rng = np.random.default_rng(10)
    t_all = 0.0
    fits = 0
    items = []
    for r in range(R):
        reg = nb.region(r)
        rs = cpn.RegStruct()
        rs.L, rs.Nl, rs.rho_l, rs.dl = reg["L"], reg["Nl"], reg["rho_l"], reg["d_l"]
        n0, n1_ = nb.cpg_off[r], nb.cpg_off[r + 1]
        rs.N, rs.rho_n, rs.dn = int(n1_ - n0), nb.rho_n[n0:n1_], nb.d_n[n0 - r:n1_ - r - 1]
        rs.cpg_pos, rs.chrst, rs.chrend = nb.cpg_pos[n0:n1_], int(nb.chrst[r]), int(nb.chrend[r])
        eng.get_nls_reg_info(rs, cfg)
        perms = np.stack([np.sort(rng.choice(np.arange(1, 2 * n1 + 1), n1, replace=False)) for _ in range(P)])
        items.append((rs, (reg["lu"][:n1], reg["lm"][:n1]), (reg["lu"][n1:], reg["lm"][n1:]), perms))
    G = args.regions_per_call
    for rep in range(2):                                                     # first pass warms the workspaces up
        t_all, fits = 0.0, 0
        for g0 in range(0, R, G):
            grp = items[g0:g0 + G]
            t0 = time.perf_counter()
            Tn = (eng.two_samp_null_stats_device if args.device_views else eng.two_samp_null_stats_regions)(grp, cfg, rng=rng)
            for (rs, _, _, _), T in zip(grp, Tn):
                K = rs.nls_rgs.num
                lib.two_samp_pvals(np.array([0, K]), P, T[0].reshape(-1), T.reshape(-1), False)
            t_all += time.perf_counter() - t0
            fits += 2 * P * len(grp)
    print(json.dumps({"workload": "two-sample permutation test, 30+30 reads", "regions": R, "perms": P, "em_fits": fits,
                      "starts_per_fit": cfg.max_em_init, "regions_per_call": G, "device_views": bool(args.device_views), "wall_s": t_all, "regions_per_s_wall": R / t_all, "fits_per_s_wall": fits / t_all,
                      "note": ("cpn_two_samp_null_batch: pooled reads uploaded once, refits are views" if args.device_views else
                               "permuted samples materialised on the host: cpn_fit_batch + cpn_stat_sums_batch + cpn_cmd_batch") +
                              "; regions_per_call regions per call; pageable host buffers; wall includes drawing the EM starts"}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", choices=["group", "two_sample"], required=True)
    ap.add_argument("--regions", type=int, default=None)
    ap.add_argument("--perms", type=int, default=1000)
    ap.add_argument("--regions-per-call", type=int, default=16)
    ap.add_argument("--device-views", type=int, default=1, help="1: cpn_two_samp_null_batch (refits are views); 0: materialise permuted reads")
    a = ap.parse_args()
    if a.regions is None:
        a.regions = 20000 if a.workload == "group" else 32
    (group if a.workload == "group" else two_sample)(a)
