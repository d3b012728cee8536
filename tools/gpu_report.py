#!/usr/bin/env python3
"""Diagnostic report run on the GPU box: parity error distributions against the oracle and per-kernel timings.
Writes gpurun_out/report_<tag>.txt.  Test/bench infrastructure — not part of the product."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r01")
    ap.add_argument("--regions", type=int, default=4096)
    ap.add_argument("--parity-regions", type=int, default=48)
    ap.add_argument("--n-init", type=int, default=10)
    args = ap.parse_args()
    import cpelnano_b200 as cpn
    import oracle_py as orc
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", f"report_{args.tag}.txt"), "w")

    def P(*a):
        s = " ".join(str(x) for x in a)
        print(s)
        out.write(s + "\n")
        out.flush()

    eng = cpn.Engine(0)
    lib = eng.lib
    P("library:", lib.version())
    P("FP64 DFMA peak (measured, TFLOP/s):", round(lib.measure_fp64_peak(), 2))

    # ---- parity distributions ------------------------------------------------------------------
    nb = cpn.simulations.make_nano_batch(args.parity_regions, M=30, seed=101)
    rng = np.random.default_rng(7)
    n_init = 4
    phi0 = cpn.simulations.gen_phi0s(rng, nb.R, n_init)
    ES, lp = lib.estep_batch(nb.grp_off, nb.Nl, nb.rho_l, nb.d_l, nb.read_off, nb.log_pyx_u, nb.log_pyx_m, phi0[:, 0])
    e_es, e_lp = [], []
    for r in range(nb.R):
        rg = nb.region(r)
        es_o = orc.e_step(rg["Nl"], rg["rho_l"], rg["d_l"], rg["lu"], rg["lm"], phi0[r, 0])
        lp_o = orc.comp_ave_log_py(rg["Nl"], rg["rho_l"], rg["d_l"], rg["lu"], rg["lm"], phi0[r, 0])
        e_es.append(np.max(np.abs(ES[r] - es_o) / np.maximum(1, np.abs(es_o))))
        e_lp.append(abs(lp[r] - lp_o) / max(1, abs(lp_o)))
    P("E-step ES   max rel err vs oracle:", np.max(e_es), " median:", np.median(e_es))
    P("ave log p(y) max rel err vs oracle:", np.max(e_lp))

    t0 = time.time()
    ref = orc.fit_batch(nb.grp_off, nb.Nl, nb.rho_l, nb.d_l, nb.read_off, nb.log_pyx_u, nb.log_pyx_m, phi0, n_init, 20, n_threads=os.cpu_count())
    t_orc = time.time() - t0
    P(f"oracle fit: {nb.R} regions x {n_init} starts in {t_orc:.2f}s on {os.cpu_count()} threads → {nb.R / t_orc:.2f} regions/s")
    ref2 = orc.fit_batch(nb.grp_off, nb.Nl, nb.rho_l, nb.d_l, nb.read_off, nb.log_pyx_u * (1 + 2.3e-16), nb.log_pyx_m * (1 + 2.3e-16), phi0,
                         n_init, 20, n_threads=os.cpu_count())
    for mode, name in ((cpn.MSTEP_ANALYTIC, "analytic"), (cpn.MSTEP_FD, "fd")):
        got = lib.fit_batch(nb.grp_off, nb.Nl, nb.rho_l, nb.d_l, nb.read_off, nb.log_pyx_u, nb.log_pyx_m, phi0, n_init, 20, mstep_mode=mode,
                            per_start=True)
        P(f"--- GPU mstep_mode={name} vs oracle, per start ---  device ms {lib.last_device_ms():.2f} launches {lib.last_launches()}")
        report_pair(P, got, ref)
    P("--- oracle vs oracle with emissions perturbed by 1 ulp (the reference's own reproducibility) ---")
    report_pair(P, ref2, ref)

    # ---- timings --------------------------------------------------------------------------------
    nb = cpn.simulations.make_nano_batch(args.regions, M=30, seed=5)
    phi0 = cpn.simulations.gen_phi0s(np.random.default_rng(1), nb.R, args.n_init)
    for mode, name in ((cpn.MSTEP_ANALYTIC, "analytic"), (cpn.MSTEP_FD, "fd")):
        if mode == cpn.MSTEP_FD and args.regions > 8192:
            continue
        lib.fit_batch(nb.grp_off, nb.Nl, nb.rho_l, nb.d_l, nb.read_off, nb.log_pyx_u, nb.log_pyx_m, phi0, args.n_init, 20, mstep_mode=mode)
        lib.set_profile(True)
        t0 = time.time()
        o = lib.fit_batch(nb.grp_off, nb.Nl, nb.rho_l, nb.d_l, nb.read_off, nb.log_pyx_u, nb.log_pyx_m, phi0, args.n_init, 20, mstep_mode=mode)
        wall = time.time() - t0
        prof = lib.get_profile()
        lib.set_profile(False)
        ms = lib.last_device_ms()
        c = o["counters"].sum(0)
        P(f"=== fit {nb.R} regions x {args.n_init} starts, mode={name}: device {ms:.1f} ms, wall {wall * 1e3:.1f} ms → {nb.R / (ms * 1e-3):.0f} regions/s (device)")
        P("    counters: em_iters/region %.1f  solver_iters/region %.1f  evals/region %.1f  converged regions %.3f" % (
            c[0] / nb.R, c[1] / nb.R, c[2] / nb.R, np.mean(o["pick"] >= 0)))
        P("    per-kernel-class ms (launches):", json.dumps({k: (round(v[0], 2), v[1]) for k, v in prof.items() if v[1]}))
    out.close()


def report_pair(P, got, ref):
    cg, co = np.isfinite(got["Q_starts"]), np.isfinite(ref["Q_starts"])
    both = cg & co
    P("   starts converged: a %.3f b %.3f agree %.3f" % (cg.mean(), co.mean(), (cg == co).mean()))
    d = np.abs(got["phi_starts"][both] - ref["phi_starts"][both])
    for q in (50, 90, 99, 100):
        P(f"   |dphi| p{q}: a={np.percentile(d[:, 0], q):.2e} b={np.percentile(d[:, 1], q):.2e} c={np.percentile(d[:, 2], q):.2e}")
    dq = np.abs(got["Q_starts"][both] - ref["Q_starts"][both]) / np.abs(ref["Q_starts"][both])
    P("   |dQ|/|Q| median %.2e max %.2e" % (np.median(dq), np.max(dq)))
    P("   frac of starts with all |dphi| <= 1e-6*max(1,|phi|): %.3f" % np.mean(np.all(d <= 1e-6 * np.maximum(1, np.abs(ref["phi_starts"][both])), axis=1)))
    ok = (got["pick"] >= 0) & (ref["pick"] >= 0)
    P("   picked: same start %.3f ; |dQhat|/|Q| max %.2e" % (np.mean(got["pick"][ok] == ref["pick"][ok]), np.max(np.abs(got["Q"][ok] - ref["Q"][ok]) / np.abs(ref["Q"][ok]))))


if __name__ == "__main__":
    main()
