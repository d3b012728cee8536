"""WGBS front-end: get_ϕhat_wgbs! / em_iter_wgbs (src/em_algorithm.jl:412-575) and pmap_analyze_reg_bam
(src/wgbs.jl:335-389) on per-CpG calls.  The BAM/XM-tag parser (wgbs.jl:222-323) is outside the library; its output format — one
call in {−1, 0, +1} per read and CpG site — is the input here.

CPU: the packing of calls into hard emissions against the reference's rule (wgbs.jl:294-304), and the oracle's restatement of the
non-log E-step em_iter_wgbs uses (get_cond_exs) against its log-domain twin on the same emissions.
GPU: E-step statistics of hard emissions against the oracle's non-log E-step; cpn_analyze_wgbs_batch against cpn_analyze_batch
fed the explicit 0 / −50 tables (bit-identical), and the host mirror analyze_regions_wgbs with the reference's coverage filters."""
import numpy as np
import pytest


def _wgbs(cpn, R=6, M=30, seed=4, pobs=0.5):
    return cpn.simulations.make_wgbs_batch(R, M=M, seed=seed, pobs=pobs)


def test_wgbs_pack_follows_the_hard_emission_rule(cpn):
    lib = cpn.Library()
    nb = _wgbs(cpn)
    llr, base = lib.wgbs_pack(nb.cpg_off, nb.read_off, nb.calls, 3)
    # wgbs.jl:294-304: x=−1 → (log_pyx_u, log_pyx_m) = (0, −50); x=+1 → (−50, 0); pack = (m − u, Σ_obs u)
    want = np.where(nb.calls == 1, 50.0, np.where(nb.calls == -1, -50.0, np.nan))
    assert np.array_equal(np.isnan(llr), nb.calls == 0) and np.array_equal(llr[nb.calls != 0], want[nb.calls != 0])
    llr2, base2 = lib.pack_emissions(nb.grp_off, nb.read_off, nb.log_pyx_u, nb.log_pyx_m)
    assert np.array_equal(llr[nb.calls != 0], llr2[nb.calls != 0]) and np.array_equal(base, base2)
    bad = nb.calls.copy(); bad[3] = 2
    with pytest.raises(cpn.capi.CpnError):
        lib.wgbs_pack(nb.cpg_off, nb.read_off, bad)


def test_wgbs_coverage_pattern_is_fragment_like(cpn):
    rng = np.random.default_rng(0)
    obs = cpn.simulations.wgbs_obs_pattern(rng, 200, 120, 0.25)
    runs = []
    for row in obs:
        k = 0
        for v in list(row) + [False]:
            if v:
                k += 1
            elif k:
                runs.append(k); k = 0
    assert max(runs) <= 9 and np.median(runs) == 9          # a fragment covers its first nine groups (simulations.jl:747-752)
    assert 0.2 < obs.mean() < 0.6


def test_oracle_nonlog_estep_equals_log_estep_on_hard_calls(orc, cpn):
    """em_iter_wgbs's E-step (non-log chain, matrices scaled by their largest entry) and em_iter's (log chain) are the same
    expectations; on short chains the non-log products keep ≥ 9 digits."""
    nb = _wgbs(cpn, R=4, M=12)
    for r in range(nb.R):
        rg = nb.region(r)
        L = min(rg["L"], 14)                                   # the unscaled e^{-25·L} products leave the double range on long chains
        sl = slice(0, L)
        phi = nb.phi_true[r] * 0.5
        a = orc.e_step_wgbs(rg["Nl"][sl], rg["rho_l"][sl], rg["d_l"][:L - 1], rg["lu"][:, sl], rg["lm"][:, sl], phi)
        b = orc.e_step(rg["Nl"][sl], rg["rho_l"][sl], rg["d_l"][:L - 1], rg["lu"][:, sl], rg["lm"][:, sl], phi)
        assert np.allclose(a, b, rtol=1e-9, atol=1e-9), (a, b)


@pytest.mark.gpu
def test_gpu_estep_on_hard_calls_vs_oracle_nonlog_chain(engine, cpn, orc):
    nb = _wgbs(cpn, R=5, M=16)
    # chains cut to 14 CpGs so that the reference's non-log arithmetic stays well conditioned
    g_off, r_off, Nl, rho, dl, lu, lm, phis = [0], [0], [], [], [], [], [], []
    regs = []
    for r in range(nb.R):
        rg = nb.region(r)
        L = 14
        regs.append(dict(Nl=rg["Nl"][:L], rho=rg["rho_l"][:L], dl=rg["d_l"][:L - 1], lu=np.ascontiguousarray(rg["lu"][:, :L]),
                         lm=np.ascontiguousarray(rg["lm"][:, :L])))
        g_off.append(g_off[-1] + L); r_off.append(r_off[-1] + rg["M"])
        phis.append(nb.phi_true[r] * 0.5)
    cat = np.concatenate
    ES, _ = engine.lib.estep_batch(np.array(g_off), cat([x["Nl"] for x in regs]), cat([x["rho"] for x in regs]), cat([x["dl"] for x in regs]),
                                   np.array(r_off), cat([x["lu"].reshape(-1) for x in regs]), cat([x["lm"].reshape(-1) for x in regs]), np.array(phis))
    for r, x in enumerate(regs):
        want = orc.e_step_wgbs(x["Nl"], x["rho"], x["dl"], x["lu"], x["lm"], phis[r])
        assert np.allclose(ES[r], want, rtol=1e-9, atol=1e-9), (ES[r], want)


@pytest.mark.gpu
def test_analyze_wgbs_batch_equals_analyze_batch_on_explicit_tables(engine, cpn):
    lib = engine.lib
    nb = _wgbs(cpn, R=40, M=40, pobs=0.6)
    n_init = 3
    phi0 = cpn.simulations.gen_phi0s(np.random.default_rng(3), nb.R, n_init)
    subs = [lib.nls_reg_info(int(nb.chrst[r]), int(nb.chrend[r]), 350, nb.cpg_pos[nb.cpg_off[r]:nb.cpg_off[r + 1]]) for r in range(nb.R)]
    sub_off = np.concatenate([[0], np.cumsum([len(s[2]) for s in subs])]).astype(np.int64)
    sub_p, sub_q = np.concatenate([s[2] for s in subs]), np.concatenate([s[3] for s in subs])
    got = lib.analyze_wgbs_batch(nb.cpg_off, nb.rho_n, nb.d_n, nb.read_off, nb.calls, phi0, n_init, sub_off, sub_p, sub_q)
    ref = lib.analyze_batch(nb.cpg_off, np.ones(int(nb.cpg_off[-1])), nb.rho_n, nb.d_n, nb.read_off, nb.log_pyx_u, nb.log_pyx_m, phi0, n_init,
                            nb.cpg_off, nb.rho_n, nb.d_n, sub_off, sub_p, sub_q)
    assert (got["pick"] >= 0).mean() > 0.5
    for k in ("phi_hat", "Q", "pick", "logZ", "ex", "exx", "mml", "nme"):
        assert np.array_equal(ref[k], got[k], equal_nan=True), k
    # the fits recover the simulated mean level (sign of E[X] follows the truth) on well covered regions
    ok = got["pick"] >= 0
    assert np.nanmedian(np.abs(got["phi_hat"][ok, 0] - nb.phi_true[ok, 0])) < 1.0


@pytest.mark.gpu
def test_host_mirror_of_pmap_analyze_reg_bam(engine, cpn):
    nb = _wgbs(cpn, R=8, M=40, pobs=0.6)
    cfg = cpn.CpelNanoConfig(min_cov=5.0, max_em_init=2, max_em_iters=20)
    regs, calls = [], []
    for r in range(nb.R):
        a, b = int(nb.cpg_off[r]), int(nb.cpg_off[r + 1])
        rs = cpn.RegStruct(chr="chrT", chrst=int(nb.chrst[r]), chrend=int(nb.chrend[r]), N=b - a, L=b - a, Nl=np.ones(b - a), rho_n=nb.rho_n[a:b],
                           rho_l=nb.rho_n[a:b], dn=nb.d_n[a - r:b - r - 1], dl=nb.d_n[a - r:b - r - 1], cpg_pos=nb.cpg_pos[a:b])
        regs.append(rs)
        M = int(nb.read_off[r + 1] - nb.read_off[r])
        calls.append(nb.calls[nb.obs_off[r]:nb.obs_off[r + 1]].reshape(M, b - a))
    calls[1] = calls[1][:2]                                   # depth below min_cov: skipped like wgbs.jl:363
    cpn.analyze_regions_wgbs(regs, calls, cfg, rng=np.random.default_rng(5), engine=engine)
    assert not regs[1].proc
    done = [rs for rs in regs if rs.proc]
    assert len(done) >= 4
    for rs in done:
        assert len(rs.phihat) == 3 and np.all(np.isfinite(rs.ex)) and len(rs.mml) == rs.nls_rgs.num
        assert np.all((rs.mml[~np.isnan(rs.mml)] >= 0) & (rs.mml[~np.isnan(rs.mml)] <= 1))
