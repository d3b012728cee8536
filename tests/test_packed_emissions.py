"""Packed emissions (cpn_pack_emissions, cpn_fit_batch_packed, cpn_analyze_batch_packed): the host-side packing against a
plain Python statement of structures.jl:121-128 → (LLR, Σ_obs log p(y|−1)), and — on the GPU — bit-identity of the packed entry
points with their two-table twins, through the single-batch path and through the chunked, pipelined path."""
import os

import numpy as np
import pytest


def _python_pack(nb):
    llr = np.full(nb.log_pyx_u.size, np.nan)
    base = np.zeros(int(nb.read_off[-1]))
    for r in range(nb.R):
        rg = nb.region(r)
        o = int(nb.obs_off[r])
        for m in range(rg["M"]):
            acc = 0.0
            for l in range(rg["L"]):
                u, v = rg["lu"][m, l], rg["lm"][m, l]
                if u == u:
                    llr[o + m * rg["L"] + l] = v - u
                    acc += u
            base[int(nb.read_off[r]) + m] = acc
    return llr, base


def test_pack_emissions_matches_the_plain_statement(cpn):
    lib = cpn.Library()
    nb = cpn.simulations.make_nano_batch(40, M=9, seed=21, pobs=0.6)
    want_llr, want_base = _python_pack(nb)
    for nt in (1, 3, 16):
        llr, base = lib.pack_emissions(nb.grp_off, nb.read_off, nb.log_pyx_u, nb.log_pyx_m, nt)
        assert np.array_equal(np.isnan(llr), np.isnan(want_llr))
        assert np.array_equal(llr[~np.isnan(llr)], want_llr[~np.isnan(want_llr)])
        assert np.array_equal(base, want_base)


def test_pack_emissions_rejects_an_observed_cell_without_a_number(cpn):
    lib = cpn.Library()
    nb = cpn.simulations.make_nano_batch(3, M=4, seed=2, pobs=1.0)
    lm = nb.log_pyx_m.copy()
    lm[5] = np.nan                      # log p(y|−1) is a number, log p(y|+1) is not: cannot be told apart from "unobserved" once packed
    with pytest.raises(cpn.capi.CpnError):
        lib.pack_emissions(nb.grp_off, nb.read_off, nb.log_pyx_u, lm)
    # empty batch and ragged shapes
    llr, base = lib.pack_emissions(np.array([0]), np.array([0]), np.zeros(0), np.zeros(0))
    assert llr.size == 0 and base.size == 0


@pytest.mark.gpu
def test_packed_entry_points_are_bit_identical_to_the_two_table_ones(engine, cpn):
    lib = engine.lib
    nb = cpn.simulations.make_nano_batch(700, M=23, seed=77, pobs=0.8)
    n_init = 3
    phi0 = cpn.simulations.gen_phi0s(np.random.default_rng(9), nb.R, n_init)
    llr, base = lib.pack_emissions(nb.grp_off, nb.read_off, nb.log_pyx_u, nb.log_pyx_m)
    ref = lib.fit_batch(nb.grp_off, nb.Nl, nb.rho_l, nb.d_l, nb.read_off, nb.log_pyx_u, nb.log_pyx_m, phi0, n_init, 20, per_start=True)
    got = lib.fit_batch_packed(nb.grp_off, nb.Nl, nb.rho_l, nb.d_l, nb.read_off, llr, base, phi0, n_init, 20, per_start=True)
    for k in ("phi_hat", "Q", "pick", "counters", "phi_starts", "Q_starts"):
        assert np.array_equal(ref[k], got[k], equal_nan=True), k
    # the chunked, pipelined host path (several chunks, worker threads) slices the per-read table by reads, not by cells
    old = {k: os.environ.get(k) for k in ("CPN_CHUNK_REGIONS", "CPN_CHUNKS")}
    os.environ["CPN_CHUNK_REGIONS"] = "100"
    os.environ["CPN_CHUNKS"] = "4"
    try:
        got2 = lib.fit_batch_packed(nb.grp_off, nb.Nl, nb.rho_l, nb.d_l, nb.read_off, llr, base, phi0, n_init, 20, per_start=True)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    for k in ("phi_hat", "Q", "pick", "phi_starts", "Q_starts"):
        assert np.array_equal(ref[k], got2[k], equal_nan=True), k


@pytest.mark.gpu
def test_packed_analyze_equals_two_table_analyze(engine, cpn):
    lib = engine.lib
    nb = cpn.simulations.make_nano_batch(120, M=17, seed=5)
    n_init = 2
    phi0 = cpn.simulations.gen_phi0s(np.random.default_rng(1), nb.R, n_init)
    subs = [lib.nls_reg_info(int(nb.chrst[r]), int(nb.chrend[r]), 350, nb.cpg_pos[nb.cpg_off[r]:nb.cpg_off[r + 1]]) for r in range(nb.R)]
    sub_off = np.concatenate([[0], np.cumsum([len(s[2]) for s in subs])]).astype(np.int64)
    sub_p, sub_q = np.concatenate([s[2] for s in subs]), np.concatenate([s[3] for s in subs])
    ref = lib.analyze_batch(nb.grp_off, nb.Nl, nb.rho_l, nb.d_l, nb.read_off, nb.log_pyx_u, nb.log_pyx_m, phi0, n_init, nb.cpg_off, nb.rho_n, nb.d_n,
                            sub_off, sub_p, sub_q)
    llr, base = lib.pack_emissions(nb.grp_off, nb.read_off, nb.log_pyx_u, nb.log_pyx_m)
    R, sN, sK = nb.R, int(nb.cpg_off[-1]), int(sub_off[-1])
    out = dict(phi_hat=np.empty((R, 3)), Q=np.empty(R), pick=np.empty(R, dtype=np.int32), counters=np.empty((R, 4), dtype=np.int64),
               logZ=np.empty(R), ex=np.empty(sN), exx=np.empty(sN - R), log_g=np.empty((sK, 4)), e_log_g1=np.empty(sK), e_log_g2=np.empty(sK),
               mml=np.empty(sK), nme=np.empty(sK))
    f64, i64 = (lambda a: np.ascontiguousarray(a, dtype=np.float64)), (lambda a: np.ascontiguousarray(a, dtype=np.int64))
    lib.analyze_batch_raw(R, i64(nb.grp_off), f64(nb.Nl), f64(nb.rho_l), f64(nb.d_l), i64(nb.read_off), llr, base, f64(phi0), n_init, 20,
                          (20.0, 150.0, 50.0), cpn.MSTEP_ANALYTIC, i64(nb.cpg_off), f64(nb.rho_n), f64(nb.d_n), sub_off, i64(sub_p), i64(sub_q),
                          out["phi_hat"], out["Q"], out["pick"], out["counters"], out["logZ"], out["ex"], out["exx"], out["log_g"],
                          out["e_log_g1"], out["e_log_g2"], out["mml"], out["nme"], False, 0, packed=True)
    for k in ("phi_hat", "Q", "pick", "logZ", "ex", "exx", "mml", "nme"):
        assert np.array_equal(ref[k], out[k], equal_nan=True), k
