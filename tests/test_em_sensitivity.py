"""The reference's EM output is not reproducible to 1e-6 — measured with the oracle alone (no GPU).

The M-step (src/em_algorithm.jl:152-163) stops at ‖U‖∞ ≤ 1e-3 on a finite-difference Jacobian of a finite-difference gradient
(absolute noise ≈ 4e-4 on entries as small as 0.05) and the EM (src/em_algorithm.jl:18-23) stops at ‖Δϕ‖₂ ≤ 3e-2, so the
last-ulp rounding of the inputs decides where inside that slack an EM instance ends.  Multiplying every emission by
(1 + 2.3e-16) — a perturbation far below anything a different BLAS `dot`, libm or summation order introduces — is enough.
This is why tests/test_gpu_parity.py compares fitted ϕ per start against this measured self-reproducibility and asserts
1e-10 only on the quantities that feed ϕ̂."""
import os

import numpy as np


def test_oracle_em_moves_under_one_ulp(orc, cpn):
    nb = cpn.simulations.make_nano_batch(10, M=30, seed=1)
    n_init = 4
    phi0 = cpn.simulations.gen_phi0s(np.random.default_rng(5), nb.R, n_init)
    nt = min(8, os.cpu_count() or 1)
    A = orc.fit_batch(nb.grp_off, nb.Nl, nb.rho_l, nb.d_l, nb.read_off, nb.log_pyx_u, nb.log_pyx_m, phi0, n_init, n_threads=nt)
    B = orc.fit_batch(nb.grp_off, nb.Nl, nb.rho_l, nb.d_l, nb.read_off, nb.log_pyx_u * (1 + 2.3e-16), nb.log_pyx_m * (1 + 2.3e-16), phi0, n_init,
                      n_threads=nt)
    both = np.isfinite(A["Q_starts"]) & np.isfinite(B["Q_starts"])
    assert both.sum() >= 10
    d = np.abs(A["phi_starts"][both] - B["phi_starts"][both])
    rel = d / np.maximum(1.0, np.abs(A["phi_starts"][both]))
    # most starts move by more than 1e-6 in at least one component …
    assert np.mean(np.any(rel > 1e-6, axis=1)) > 0.5
    # … but stay inside the EM stopping slack, and the objective is unaffected to 1e-5
    assert np.median(d, axis=0).max() < 3e-2
    dq = np.abs(A["Q_starts"][both] - B["Q_starts"][both]) / np.abs(A["Q_starts"][both])
    assert np.median(dq) < 1e-6 and np.max(dq) < 1e-3
    # the perturbation itself is invisible at the 1e-10 level in what feeds the M-step
    rg = nb.region(0)
    es0 = orc.e_step(rg["Nl"], rg["rho_l"], rg["d_l"], rg["lu"], rg["lm"], phi0[0, 0])
    es1 = orc.e_step(rg["Nl"], rg["rho_l"], rg["d_l"], rg["lu"] * (1 + 2.3e-16), rg["lm"] * (1 + 2.3e-16), phi0[0, 0])
    assert np.all(np.abs(es0 - es1) <= 1e-11 * np.maximum(1, np.abs(es0)))
