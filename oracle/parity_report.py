"""Output-level EM parity report.  TEST INFRASTRUCTURE (used by tests/ and by bench.py's cpu_baseline leg only).

Fitted ϕ̂ cannot be compared at 1e-6 between ANY two implementations of the reference's EM (its M-step stops at ‖U‖∞ ≤ 1e-3 on
a finite-difference Jacobian of a finite-difference gradient and its EM at ‖Δϕ‖₂ ≤ 3e-2: tests/test_em_sensitivity.py), so
parity is stated where users see it: on the rows of the MML / NME output files, which the reference writes rounded to four
decimals (src/input_output.jl:173,217).  For two sets of fits of the same regions this module counts

    row flip rate      fraction of (region, sub-region) rows whose 4-dp MML or NME text differs, over regions both fitted
    status flip rate   fraction of regions fitted by one implementation and not by the other (rs.ϕhat == [] in one only)
    pick flip rate     fraction of regions, fitted by both, where pick_ϕ (src/em_algorithm.jl:35-53) chose a different start

and the yardstick is the same three numbers for the oracle against ITSELF with every emission multiplied by (1 + 2.3e-16) —
the reference algorithm's own reproducibility under a one-ulp change of its inputs.
"""
import os

import numpy as np


def round4(x):
    """round(x, digits=4) of Julia: half-even on x·10⁴ (Base/floatfuncs.jl)."""
    return np.rint(np.asarray(x, dtype=np.float64) * 1e4) / 1e4


def oracle_fit(orc, nb, phi0, n_init, max_em_iters=20, perturb=False, n_threads=None):
    s = (1 + 2.3e-16) if perturb else 1.0
    return orc.fit_batch(nb.grp_off, nb.Nl, nb.rho_l, nb.d_l, nb.read_off, nb.log_pyx_u * s, nb.log_pyx_m * s, phi0, n_init, max_em_iters,
                         n_threads=n_threads or os.cpu_count() or 1)


def oracle_rows(orc, nb, sub_off, sub_p, sub_q, phi_hat, pick):
    """MML / NME of every sub-region from the oracle's get_stat_sums (NaN rows where the region has no fit)."""
    sK = int(sub_off[-1])
    mml, nme = np.full(sK, np.nan), np.full(sK, np.nan)
    for r in range(nb.R):
        if pick[r] < 0:
            continue
        n0, n1 = nb.cpg_off[r], nb.cpg_off[r + 1]
        k0, k1 = sub_off[r], sub_off[r + 1]
        ss = orc.get_stat_sums(phi_hat[r], nb.rho_n[n0:n1], nb.d_n[n0 - r:n1 - r - 1], sub_p[k0:k1], sub_q[k0:k1])
        mml[k0:k1], nme[k0:k1] = ss["mml"], ss["nme"]
    return mml, nme


def compare(pick_a, mml_a, nme_a, pick_b, mml_b, nme_b, sub_off):
    """The three flip rates between two sets of fits (a, b) of the same regions."""
    ok_a, ok_b = np.asarray(pick_a) >= 0, np.asarray(pick_b) >= 0
    both = ok_a & ok_b
    K = np.diff(sub_off)
    rows_both = np.repeat(both, K)
    ra, rb = np.stack([round4(mml_a), round4(nme_a)]), np.stack([round4(mml_b), round4(nme_b)])
    have = rows_both & ~np.isnan(ra[0]) & ~np.isnan(rb[0])                   # empty sub-regions are not written (NaN rows)
    differ = np.any(ra[:, have] != rb[:, have], axis=0)
    dm = np.abs(np.asarray(mml_a)[have] - np.asarray(mml_b)[have]); dn = np.abs(np.asarray(nme_a)[have] - np.asarray(nme_b)[have])
    return {"regions": int(len(ok_a)), "regions_both_fitted": int(both.sum()), "rows": int(have.sum()),
            "row_flip_rate": float(differ.mean()) if have.any() else 0.0,
            "status_flip_rate": float(np.mean(ok_a != ok_b)),
            "pick_flip_rate": float(np.mean(np.asarray(pick_a)[both] != np.asarray(pick_b)[both])) if both.any() else 0.0,
            "median_abs_dmml": float(np.median(dm)) if have.any() else 0.0, "p99_abs_dmml": float(np.percentile(dm, 99)) if have.any() else 0.0,
            "median_abs_dnme": float(np.median(dn)) if have.any() else 0.0, "p99_abs_dnme": float(np.percentile(dn, 99)) if have.any() else 0.0}
