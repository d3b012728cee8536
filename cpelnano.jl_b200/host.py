"""Host-side mirror of the reference's Julia interface for the estimation hot path.

Everything numeric is done by libcpelnano_b200.so; this module only packs `RegStruct`s into the CSR
batches of include/cpelnano_b200.h and unpacks results, keeping the reference's names, argument meaning
and failure behaviour (empty ϕ̂ ⇒ region skipped; NaN for empty sub-regions).
"""
import math
from dataclasses import dataclass, field
from itertools import combinations

import numpy as np

from . import sharding
from .capi import Library, MSTEP_ANALYTIC

LMAX = 1000                      # src/CpelNano.jl:50
AMAX, BMAX, CMAX = 20.0, 150.0, 50.0   # src/CpelNano.jl:36-38 (non-const globals in the reference → passed by value)


@dataclass
class CpelNanoConfig:
    """src/structures.jl:36-69 (fields used by the hot path; defaults of CpelNanoConfig())."""
    min_cov: float = 10.0
    min_grp_dist: int = 10
    max_size_subreg: int = 350
    size_est_reg: int = 3000
    max_em_init: int = 10
    max_em_iters: int = 20
    matched: bool = False
    verbose: bool = False
    LMAX_TWO_SAMP: int = 100
    pval_comp: bool = True
    trim: tuple = (0, 0, 0, 0)            # WGBS: bases trimmed from the XM strings (R1 5', R1 3', R2 5', R2 3'), structures.jl:53
    pe: bool = False                      # WGBS: paired-end BAM, structures.jl:54
    mstep_mode: int = MSTEP_ANALYTIC      # B200 extension: see include/cpelnano_b200.h

    @classmethod
    def five_arg(cls, min_cov, max_size_subreg, size_est_reg, max_em_init, max_em_iters):
        """The 5-argument constructor silently stores max_size_subreg + 1 (src/structures.jl:65-68)."""
        return cls(min_cov=min_cov, max_size_subreg=max_size_subreg + 1, size_est_reg=size_est_reg, max_em_init=max_em_init,
                   max_em_iters=max_em_iters)


@dataclass
class AnalysisRegions:
    """src/structures.jl:140-147 — closed genomic intervals and 1-based CpG index ranges (0 = empty)."""
    num: int = 0
    chr_st: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int64))
    chr_end: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int64))
    cpg_p: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int64))
    cpg_q: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int64))


def _empty():
    return np.zeros(0)


@dataclass
class RegStruct:
    """src/structures.jl:150-197.  `calls` is stored packed: lu/lm [m, L] with NaN = not observed."""
    proc: bool = False
    chr: str = ""
    chrst: int = 0
    chrend: int = 0
    N: int = 0
    L: int = 0
    Nl: np.ndarray = field(default_factory=_empty)
    rho_n: np.ndarray = field(default_factory=_empty)
    rho_l: np.ndarray = field(default_factory=_empty)
    dn: np.ndarray = field(default_factory=_empty)
    dl: np.ndarray = field(default_factory=_empty)
    cpg_pos: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int64))
    m: int = 0
    log_pyx_u: np.ndarray = field(default_factory=lambda: np.zeros((0, 0)))
    log_pyx_m: np.ndarray = field(default_factory=lambda: np.zeros((0, 0)))
    nls_rgs: AnalysisRegions = field(default_factory=AnalysisRegions)
    phihat: np.ndarray = field(default_factory=_empty)
    logZ: float = math.nan
    mml: np.ndarray = field(default_factory=_empty)
    nme: np.ndarray = field(default_factory=_empty)
    ex: np.ndarray = field(default_factory=_empty)
    exx: np.ndarray = field(default_factory=_empty)
    log_g1: np.ndarray = field(default_factory=_empty)       # E[log g1(X_p)]
    log_g2: np.ndarray = field(default_factory=_empty)       # E[log g2(X_q)]
    log_gs: np.ndarray = field(default_factory=lambda: np.zeros((0, 4)))   # pm, pp, qm, qp
    Q: float = -math.inf

    group_view: object = None     # get_grp_info!-scale copy (N_l, ρ_l, d_l) kept when the struct is rescaled to the per-CpG chain

    def set_calls(self, lu, lm):
        self.log_pyx_u = np.ascontiguousarray(lu, dtype=np.float64)
        self.log_pyx_m = np.ascontiguousarray(lm, dtype=np.float64)
        self.m = self.log_pyx_u.shape[0]

    # src/structures.jl:200-203
    def get_ave_depth(self):
        return float(np.sum(~np.isnan(self.log_pyx_u))) / self.L

    def perc_gprs_obs(self):
        return float(np.sum(np.any(~np.isnan(self.log_pyx_u), axis=0))) / self.L


class Engine:
    """One cpn_ctx on one GPU."""

    def __init__(self, device=0, library=None):
        self.lib = library or Library()
        self.device = device
        self.lib.ctx(device)          # fail now if there is no B200

    # ---- packing -------------------------------------------------------------------------------
    @staticmethod
    def pack_groups(regs):
        grp_off = np.concatenate([[0], np.cumsum([r.L for r in regs])]).astype(np.int64)
        Nl = np.concatenate([np.asarray(r.Nl, dtype=np.float64) for r in regs])
        rho_l = np.concatenate([np.asarray(r.rho_l, dtype=np.float64) for r in regs])
        d_l = np.concatenate([np.asarray(r.dl, dtype=np.float64) for r in regs])
        return grp_off, Nl, rho_l, d_l

    @staticmethod
    def pack_calls(regs):
        read_off = np.concatenate([[0], np.cumsum([r.m for r in regs])]).astype(np.int64)
        lu = np.concatenate([r.log_pyx_u.reshape(-1) for r in regs])
        lm = np.concatenate([r.log_pyx_m.reshape(-1) for r in regs])
        return read_off, lu, lm

    @staticmethod
    def pack_cpgs(regs):
        cpg_off = np.concatenate([[0], np.cumsum([r.N for r in regs])]).astype(np.int64)
        rho_n = np.concatenate([np.asarray(r.rho_n, dtype=np.float64) for r in regs])
        d_n = np.concatenate([np.asarray(r.dn, dtype=np.float64) for r in regs])
        sub_off = np.concatenate([[0], np.cumsum([r.nls_rgs.num for r in regs])]).astype(np.int64)
        sub_p = np.concatenate([r.nls_rgs.cpg_p for r in regs]).astype(np.int64)
        sub_q = np.concatenate([r.nls_rgs.cpg_q for r in regs]).astype(np.int64)
        return cpg_off, rho_n, d_n, sub_off, sub_p, sub_q

    # ---- get_ϕhat! -------------------------------------------------------------------------------
    def get_phihat_batch(self, regs, config, phi0s=None, rng=None, per_start=False):
        """get_ϕhat!(rs, config) for many regions in one call (src/em_algorithm.jl:371-390).
        `phi0s` [R, max_em_init, 3] replaces gen_ϕ0s (:65-78); if None they are drawn from `rng`."""
        if not regs:
            return None
        if phi0s is None:
            from .simulations import gen_phi0s
            phi0s = gen_phi0s(rng or np.random.default_rng(), len(regs), config.max_em_init)
        grp_off, Nl, rho_l, d_l = self.pack_groups(regs)
        read_off, lu, lm = self.pack_calls(regs)
        out = self.lib.fit_batch(grp_off, Nl, rho_l, d_l, read_off, lu, lm, phi0s, config.max_em_init, config.max_em_iters,
                                 (AMAX, BMAX, CMAX), config.mstep_mode, per_start=per_start, device=self.device)
        for i, rs in enumerate(regs):
            ok = out["pick"][i] >= 0
            rs.phihat = out["phi_hat"][i].copy() if ok else np.zeros(0)     # rs.ϕhat = [] when Qhat == -Inf (:385)
            rs.Q = float(out["Q"][i])
        return out

    # ---- get_nls_reg_info! + get_stat_sums! ---------------------------------------------------------
    def get_nls_reg_info(self, rs, config):
        st, en, p, q = self.lib.nls_reg_info(rs.chrst, rs.chrend, config.max_size_subreg, rs.cpg_pos)
        rs.nls_rgs = AnalysisRegions(len(st), st, en, p, q)

    def get_stat_sums_batch(self, regs, config):
        """get_stat_sums!(rs, config) (src/statistical_summaries.jl:470-500) for regions that were rescaled with
        rscle_grp_mod (per-CpG chain) and have a non-empty ϕ̂."""
        if not regs:
            return
        for rs in regs:
            assert len(rs.phihat) == 3, "get_stat_sums needs a fitted ϕ̂"
            self.get_nls_reg_info(rs, config)
        cpg_off, rho_n, d_n, sub_off, sub_p, sub_q = self.pack_cpgs(regs)
        phi = np.array([rs.phihat for rs in regs])
        out = self.lib.stat_sums_batch(phi, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, device=self.device)
        for i, rs in enumerate(regs):
            n0, n1, k0, k1 = cpg_off[i], cpg_off[i + 1], sub_off[i], sub_off[i + 1]
            rs.logZ = float(out["logZ"][i])
            rs.ex = out["ex"][n0:n1].copy(); rs.exx = out["exx"][n0 - i:n1 - i - 1].copy()
            rs.log_gs = out["log_g"][k0:k1].copy()
            rs.log_g1 = out["e_log_g1"][k0:k1].copy(); rs.log_g2 = out["e_log_g2"][k0:k1].copy()
            rs.mml = out["mml"][k0:k1].copy(); rs.nme = out["nme"][k0:k1].copy()

    # ---- comp_cmd ------------------------------------------------------------------------------------
    def comp_cmd_pairs(self, models, pairs, regions=None):
        """comp_cmd(models[a], models[b]) for every (a,b) in pairs (src/statistical_summaries.jl:512-570).
        All models must already hold get_stat_sums outputs.  Models of a pair must share the region features."""
        if not pairs:
            return []
        # unique genomic regions by identity of the feature arrays
        key = {}
        reg_of = np.zeros(len(models), dtype=np.int64)
        reps = []
        for t, m in enumerate(models):
            kk = (m.chr, m.chrst, m.chrend, m.N)
            if kk not in key:
                key[kk] = len(reps); reps.append(m)
            reg_of[t] = key[kk]
        cpg_off, rho_n, d_n, sub_off, sub_p, sub_q = self.pack_cpgs(reps)
        phi_tbl = np.array([m.phihat for m in models])
        ex_off = np.concatenate([[0], np.cumsum([m.N for m in models])]).astype(np.int64)
        nme_off = np.concatenate([[0], np.cumsum([m.nls_rgs.num for m in models])]).astype(np.int64)
        ex_tbl = np.concatenate([m.ex for m in models]); exx_tbl = np.concatenate([m.exx for m in models])
        nme_tbl = np.concatenate([m.nme for m in models])
        pa = np.array([p[0] for p in pairs], dtype=np.int64); pb = np.array([p[1] for p in pairs], dtype=np.int64)
        cmd, cmd_off = self.lib.cmd_batch(pa, pb, phi_tbl, reg_of, ex_off, ex_tbl, exx_tbl, nme_off, nme_tbl, cpg_off, rho_n, d_n, sub_off,
                                          sub_p, sub_q, device=self.device)
        return [cmd[cmd_off[i]:cmd_off[i + 1]].copy() for i in range(len(pairs))]

    # ---- group tests -----------------------------------------------------------------------------------
    def unmat_est_reg_test_batch(self, regions_g1, regions_g2, labelings=None, rng=None):
        """unmat_est_reg_test(ms_g1, ms_g2) (src/grp_perm_tests.jl:200-291) for a list of regions; regions_g1[r] and
        regions_g2[r] are the lists of fitted models of the two groups (same n1, n2 for every region of a call)."""
        R = len(regions_g1)
        n1, n2 = len(regions_g1[0]), len(regions_g2[0])
        S = n1 + n2
        L = math.comb(S, n1)
        if L > 20:
            exact = L < LMAX
            if labelings is None:   # all of them, or rand(1:L, LMAX) de-duplicated (:242-245), unranked instead of enumerated
                labelings, exact = sharding.labelings_for(S, n1, LMAX, rng)
        else:
            exact, labelings = True, np.zeros((0, n1), dtype=np.int32)
        models, pairs, tbl_off = [], [], [0]
        for r in range(R):
            base = len(models)
            models.extend(regions_g1[r]); models.extend(regions_g2[r])
            pairs.extend((base + i, base + j) for i in range(S) for j in range(i + 1, S))
            tbl_off.append(tbl_off[-1] + regions_g1[r][0].nls_rgs.num)
        cmds = self.comp_cmd_pairs(models, pairs)
        mml = np.concatenate([np.stack([m.mml for m in regions_g1[r] + regions_g2[r]]).reshape(-1) for r in range(R)])
        nme = np.concatenate([np.stack([m.nme for m in regions_g1[r] + regions_g2[r]]).reshape(-1) for r in range(R)])
        tbl = []
        ci = 0
        for r in range(R):
            K = regions_g1[r][0].nls_rgs.num
            t = np.full((S, S, K), np.nan)
            for i in range(S):
                for j in range(i + 1, S):
                    t[i, j] = cmds[ci]; ci += 1
            tbl.append(t.reshape(-1))
        T, cnt, p = self.lib.grp_unmat_sweep(n1, n2, np.array(tbl_off), mml, nme, np.concatenate(tbl), labelings, exact, device=self.device)
        out = []
        for r in range(R):
            K = tbl_off[r + 1] - tbl_off[r]
            sl = slice(3 * tbl_off[r], 3 * tbl_off[r + 1])
            out.append(dict(T=T[sl].reshape(3, K), cnt=cnt[sl].reshape(3, K), pval=p[sl].reshape(3, K), coords=regions_g1[r][0].nls_rgs))
        return out

    def grp_comp_batch(self, regions_g1, regions_g2, matched=False, labelings=None, js=None, rng=None):
        """The body of the region loop of the group comparison (grp_perm_tests.jl:599 / :618) for a list of regions in one
        device call: the models only need ϕ̂ and the region features — get_stat_sums!, comp_cmd for every pair and the
        labeling sweep all run inside cpn_grp_test_batch.  Returns the same dicts as unmat_/mat_est_reg_test_batch."""
        R = len(regions_g1)
        n1, n2 = len(regions_g1[0]), len(regions_g2[0])
        S = n1 + n2
        reps = [regions_g1[r][0] for r in range(R)]
        cpg_off, rho_n, d_n, sub_off, sub_p, sub_q = self.pack_cpgs(reps)
        phi = np.array([[m.phihat for m in regions_g1[r] + regions_g2[r]] for r in range(R)], dtype=np.float64)
        if matched:
            if 0.5 ** n1 < 0.05:
                exact = 2 ** n1 < LMAX
                if js is None:
                    js = np.arange(2 ** n1, dtype=np.int64) if exact else (rng or np.random.default_rng()).integers(0, 2 ** n1, size=LMAX)
            else:
                exact, js = True, np.zeros(0, dtype=np.int64)
            T, cnt, p, tcmd = self.lib.grp_test_batch(n1, n2, True, phi, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, js, exact, device=self.device)
            return [dict(T=T[2 * sub_off[r]:2 * sub_off[r + 1]].reshape(2, -1), cnt=cnt[2 * sub_off[r]:2 * sub_off[r + 1]].reshape(2, -1),
                         pval=p[2 * sub_off[r]:2 * sub_off[r + 1]].reshape(2, -1), tcmd=tcmd[sub_off[r]:sub_off[r + 1]], coords=reps[r].nls_rgs)
                    for r in range(R)]
        L = math.comb(S, n1)
        if L > 20:
            exact = L < LMAX
            if labelings is None:
                labelings, exact = sharding.labelings_for(S, n1, LMAX, rng)
        else:
            exact, labelings = True, np.zeros((0, n1), dtype=np.int32)
        T, cnt, p = self.lib.grp_test_batch(n1, n2, False, phi, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, labelings, exact, device=self.device)
        return [dict(T=T[3 * sub_off[r]:3 * sub_off[r + 1]].reshape(3, -1), cnt=cnt[3 * sub_off[r]:3 * sub_off[r + 1]].reshape(3, -1),
                     pval=p[3 * sub_off[r]:3 * sub_off[r + 1]].reshape(3, -1), coords=reps[r].nls_rgs) for r in range(R)]

    def mat_est_reg_test_batch(self, regions_g1, regions_g2, js=None, rng=None):
        """mat_est_reg_test (src/grp_perm_tests.jl:432-499): Tmml, Tnme with sign-flip p-values; Tcmd = mean pair CMD, no p-value."""
        R = len(regions_g1)
        n = len(regions_g1[0])
        if 0.5 ** n < 0.05:
            exact = 2 ** n < LMAX
            if js is None:
                js = np.arange(2 ** n, dtype=np.int64) if exact else (rng or np.random.default_rng()).integers(0, 2 ** n, size=LMAX)
        else:
            exact, js = True, np.zeros(0, dtype=np.int64)
        tbl_off = np.concatenate([[0], np.cumsum([regions_g1[r][0].nls_rgs.num for r in range(R)])]).astype(np.int64)
        cat = lambda regs, f: np.concatenate([np.stack([getattr(m, f) for m in regs[r]]).reshape(-1) for r in range(R)])   # noqa: E731
        T, cnt, p = self.lib.grp_mat_sweep(n, tbl_off, cat(regions_g1, "mml"), cat(regions_g2, "mml"), cat(regions_g1, "nme"),
                                           cat(regions_g2, "nme"), js, exact, device=self.device)
        models, pairs = [], []
        for r in range(R):
            base = len(models)
            models.extend(regions_g1[r]); models.extend(regions_g2[r])
            pairs.extend((base + s, base + n + s) for s in range(n))
        cmds = self.comp_cmd_pairs(models, pairs)
        out = []
        for r in range(R):
            K = tbl_off[r + 1] - tbl_off[r]
            sl = slice(2 * tbl_off[r], 2 * tbl_off[r + 1])
            pc = cmds[r * n:(r + 1) * n]
            tc = pc[0].copy()
            for s in range(1, n):
                tc = tc + pc[s]                                        # sum(cmds)/n, left to right (:308-314)
            out.append(dict(T=T[sl].reshape(2, K), cnt=cnt[sl].reshape(2, K), pval=p[sl].reshape(2, K), tcmd=tc / n,
                            coords=regions_g1[r][0].nls_rgs))
        return out


    # ---- two-sample test ---------------------------------------------------------------------------------
    def two_samp_null_stats_batch(self, rs_ref, calls_s1, calls_s2, perms, config, phi0s=None, rng=None):
        """pmap_two_samp_null_stats (src/two_samp_perm_tests.jl:41-87) for ALL permutations of a region in one batch.
        perms [P, n1]: ascending 1-based indices into vcat(calls_s1, calls_s2) that form sample 1; the rest, in order, is
        sample 2.  Every (permutation, half) becomes one row of a CSR batch: 2P multi-start EM fits, 2P summaries, P CMDs.
        phi0s [P, 2, max_em_init, 3] replaces the 2P gen_ϕ0s draws.  Returns T_null [P, 3, K] (NaN where a fit failed)."""
        return self.two_samp_null_stats_regions([(rs_ref, calls_s1, calls_s2, perms)], config, None if phi0s is None else [phi0s], rng)[0]

    def two_samp_null_stats_regions(self, items, config, phi0s=None, rng=None):
        """The same for several regions in ONE device batch (the pmap over regions of two_samp_perm_tests.jl:220): items is a
        list of (rs_ref, calls_s1, calls_s2, perms); row order = region, permutation, half.  A region contributes only
        2·P fits, which alone cannot fill a B200 — batching regions is what brings the throughput of the EM kernels.
        Returns one T_null [P_r, 3, K_r] per region."""
        lus, lms, Nls, rhos, dls, reads, grp_len, reg_rows = [], [], [], [], [], [], [], []
        for (rs_ref, calls_s1, calls_s2, perms) in items:
            perms = np.asarray(perms, dtype=np.int64)
            P, n1 = perms.shape
            lu = np.concatenate([calls_s1[0], calls_s2[0]]); lm = np.concatenate([calls_s1[1], calls_s2[1]])
            n, L = lu.shape
            n2 = n - n1
            mask = np.ones((P, n), dtype=bool)
            mask[np.arange(P)[:, None], perms - 1] = False
            rest = np.nonzero(mask)[1].reshape(P, n2)                      # deleteat!(aux_calls, perm_ids): the rest, in order
            order = np.concatenate([perms - 1, rest], axis=1)              # [P, n]: rows of half 1 then rows of half 2
            lus.append(lu[order].reshape(-1)); lms.append(lm[order].reshape(-1))
            R2 = 2 * P
            Nls.append(np.tile(np.asarray(rs_ref.Nl, dtype=np.float64), R2)); rhos.append(np.tile(np.asarray(rs_ref.rho_l, dtype=np.float64), R2))
            dls.append(np.tile(np.asarray(rs_ref.dl, dtype=np.float64), R2))
            reads.append(np.tile(np.array([n1, n2], dtype=np.int64), P))
            grp_len.append(np.full(R2, L, dtype=np.int64))
            reg_rows.append(R2)
        rows = int(sum(reg_rows))
        grp_off = np.concatenate([[0], np.cumsum(np.concatenate(grp_len))]).astype(np.int64)
        read_off = np.concatenate([[0], np.cumsum(np.concatenate(reads))]).astype(np.int64)
        if phi0s is None:
            from .simulations import gen_phi0s
            phi0 = gen_phi0s(rng or np.random.default_rng(), rows, config.max_em_init)
        else:
            phi0 = np.concatenate([np.asarray(p).reshape(-1, config.max_em_init, 3) for p in phi0s])
        fit = self.lib.fit_batch(grp_off, np.concatenate(Nls), np.concatenate(rhos), np.concatenate(dls), read_off, np.concatenate(lus),
                                 np.concatenate(lms), phi0.reshape(rows, config.max_em_init, 3), config.max_em_init, config.max_em_iters,
                                 (AMAX, BMAX, CMAX), config.mstep_mode, device=self.device)
        refs = [it[0] for it in items]
        cpg_off, rho_n, d_n, sub_off, sub_p, sub_q = self.pack_cpgs(refs)
        reg_of = np.repeat(np.arange(len(items), dtype=np.int64), reg_rows)
        ss = self.lib.stat_sums_batch(fit["phi_hat"], cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, reg_of=reg_of, device=self.device)
        pa = np.arange(0, rows, 2, dtype=np.int64)
        cmd, cmd_off = self.lib.cmd_batch(pa, pa + 1, fit["phi_hat"], reg_of, ss["ex_off"], ss["ex"], ss["exx"], ss["k_off"], ss["nme"], cpg_off,
                                          rho_n, d_n, sub_off, sub_p, sub_q, device=self.device)
        out, row0, pair0 = [], 0, 0
        for i, (rs_ref, _, _, perms) in enumerate(items):
            P, K = len(perms), rs_ref.nls_rgs.num
            k0 = ss["k_off"][row0]
            mml = ss["mml"][k0:k0 + 2 * P * K].reshape(P, 2, K); nme = ss["nme"][k0:k0 + 2 * P * K].reshape(P, 2, K)
            c0 = cmd_off[pair0]
            T = np.stack([mml[:, 0] - mml[:, 1], nme[:, 0] - nme[:, 1], cmd[c0:c0 + P * K].reshape(P, K)], axis=1)
            bad = (fit["pick"][row0:row0 + 2 * P].reshape(P, 2) < 0).any(axis=1)     # either ϕ̂ empty → all three statistics NaN (:65)
            T[bad] = np.nan
            out.append(T)
            row0 += 2 * P; pair0 += P
        return out

    def two_samp_null_stats_device(self, items, config, phi0s=None, rng=None):
        """Same contract as two_samp_null_stats_regions through cpn_two_samp_null_batch: the pooled reads of each region are
        uploaded once and the 2·P refits are views of them (no per-permutation copies on the host or over PCIe); summaries,
        CMD and the three statistics stay on the device.  Bit-identical to the materialised path."""
        refs = [it[0] for it in items]
        grp_off = np.concatenate([[0], np.cumsum([len(r.Nl) for r in refs])]).astype(np.int64)
        Nl = np.concatenate([np.asarray(r.Nl, dtype=np.float64) for r in refs]); rho_l = np.concatenate([np.asarray(r.rho_l, dtype=np.float64) for r in refs])
        d_l = np.concatenate([np.asarray(r.dl, dtype=np.float64) for r in refs])
        lu = np.concatenate([np.concatenate([it[1][0], it[2][0]]).reshape(-1) for it in items])
        lm = np.concatenate([np.concatenate([it[1][1], it[2][1]]).reshape(-1) for it in items])
        read_off = np.concatenate([[0], np.cumsum([it[1][0].shape[0] + it[2][0].shape[0] for it in items])]).astype(np.int64)
        perms = [np.asarray(it[3], dtype=np.int32) for it in items]
        perm_off = np.concatenate([[0], np.cumsum([p.shape[0] for p in perms])]).astype(np.int64)
        n1 = np.array([p.shape[1] for p in perms], dtype=np.int64)
        rows = 2 * int(perm_off[-1])
        if phi0s is None:
            from .simulations import gen_phi0s
            phi0 = gen_phi0s(rng or np.random.default_rng(), rows, config.max_em_init)
        else:
            phi0 = np.concatenate([np.asarray(p).reshape(-1, config.max_em_init, 3) for p in phi0s])
        cpg_off, rho_n, d_n, sub_off, sub_p, sub_q = self.pack_cpgs(refs)
        T, status = self.lib.two_samp_null_batch(grp_off, Nl, rho_l, d_l, read_off, lu, lm, perm_off, n1, np.concatenate([p.reshape(-1) for p in perms]),
                                                 phi0, config.max_em_init, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, config.max_em_iters,
                                                 (AMAX, BMAX, CMAX), config.mstep_mode, device=self.device)
        out, t0 = [], 0
        for r, p in zip(refs, perms):
            n = p.shape[0] * 3 * r.nls_rgs.num
            out.append(T[t0:t0 + n].reshape(p.shape[0], 3, r.nls_rgs.num))
            t0 += n
        return out

    def diff_two_samp_comp_region(self, mod_s1, mod_s2, calls_s1, calls_s2, config, rng=None, perms=None, phi0s=None, seed=None):
        """pmap_diff_two_samp_comp (src/two_samp_perm_tests.jl:157-273) for one region: observed statistics from the two fitted
        models, null statistics from refits of permuted reads, p-values by counting.  Returns T_null too."""
        return self.diff_two_samp_comp_batch([(mod_s1, mod_s2, calls_s1, calls_s2)], config, rng=rng, perms=None if perms is None else [perms],
                                             phi0s=None if phi0s is None else [phi0s], seed=seed, want_null=True)[0]

    def diff_two_samp_comp_batch(self, items, config, rng=None, perms=None, phi0s=None, seed=None, want_null=False):
        """pmap_diff_two_samp_comp for many regions through cpn_two_samp_test_batch: items = [(mod_s1, mod_s2, calls_s1, calls_s2)],
        models with ϕ̂ and the per-CpG features (summaries are recomputed on the device).  Observed statistics, the null refits of
        ALL regions, NaN filter and counting happen in one device call per group of regions; only (T, cnt, n_valid, p) come back
        (plus T_null when want_null).  Permutations: `perms[i]` if given; all C(n, n1) splits when fewer than
        config.LMAX_TWO_SAMP (exact, :195); else LMAX_TWO_SAMP sorted samples (:211-214) drawn on the host from `rng` — or, when
        `seed` is given, drawn on the device together with the EM starts (nothing random crosses PCIe)."""
        import copy
        R = len(items)
        if R == 0:
            return []
        outs = [None] * R
        todo = {False: [], True: []}            # regions whose index sets are explicit / drawn on the device
        for i, (m1, m2, c1, c2) in enumerate(items):
            n1, n2 = c1[0].shape[0], c2[0].shape[0]
            P, exact, pr = 0, False, None
            if config.pval_comp and n1 > 0 and n2 > 0:
                Lc = math.comb(n1 + n2, n1)
                if Lc > 20:                                                   # :192
                    exact = Lc < config.LMAX_TWO_SAMP
                    if perms is not None and perms[i] is not None:
                        pr = np.asarray(perms[i], dtype=np.int32)
                    elif exact:
                        pr = np.array(list(combinations(range(1, n1 + n2 + 1), n1)), dtype=np.int32)
                    elif seed is None:    # sort!(StatsBase.sample(1:n, n1; replace=false)) LMAX_TWO_SAMP times (:211-214)
                        rng = rng or np.random.default_rng()
                        pr = np.stack([np.sort(rng.choice(np.arange(1, n1 + n2 + 1), n1, replace=False)) for _ in range(config.LMAX_TWO_SAMP)]).astype(np.int32)
                    P = config.LMAX_TWO_SAMP if pr is None else len(pr)
            todo[pr is None and P > 0].append((i, P, exact, pr))
        for on_device, grp in todo.items():
            if not grp:
                continue
            idx = [g[0] for g in grp]
            refs = []
            for i in idx:
                m1 = items[i][0]
                ref = copy.copy(m1.group_view if m1.group_view is not None else m1)    # "back to group scale" (:198-199)
                ref.nls_rgs, ref.rho_n, ref.dn, ref.N = m1.nls_rgs, m1.rho_n, m1.dn, m1.N
                if ref.nls_rgs.num == 0:
                    self.get_nls_reg_info(ref, config)
                refs.append(ref)
            has_reads = [items[i][2][0].shape[0] + items[i][3][0].shape[0] > 0 for i in idx]
            # regions without reads (pval_comp off) still need a well-formed pooled sample for the ABI: one blank read per sample
            pooled = []
            for i, ref, hr in zip(idx, refs, has_reads):
                if hr:
                    pooled.append((np.concatenate([items[i][2][0], items[i][3][0]]), np.concatenate([items[i][2][1], items[i][3][1]]), items[i][2][0].shape[0]))
                else:
                    blank = np.full((2, len(ref.Nl)), np.nan)
                    pooled.append((blank, blank, 1))
            grp_off = np.concatenate([[0], np.cumsum([len(r.Nl) for r in refs])]).astype(np.int64)
            Nl = np.concatenate([np.asarray(r.Nl, dtype=np.float64) for r in refs]); rho_l = np.concatenate([np.asarray(r.rho_l, dtype=np.float64) for r in refs])
            d_l = np.concatenate([np.asarray(r.dl, dtype=np.float64) for r in refs])
            lu = np.concatenate([p[0].reshape(-1) for p in pooled]); lm = np.concatenate([p[1].reshape(-1) for p in pooled])
            read_off = np.concatenate([[0], np.cumsum([p[0].shape[0] for p in pooled])]).astype(np.int64)
            n1 = np.array([p[2] for p in pooled], dtype=np.int64)
            perm_off = np.concatenate([[0], np.cumsum([g[1] for g in grp])]).astype(np.int64)
            exact = np.array([int(g[2]) for g in grp], dtype=np.int32)
            perm_ids = None if on_device else (np.concatenate([g[3].reshape(-1) for g in grp if g[1] > 0]) if perm_off[-1] > 0 else np.zeros(0, dtype=np.int32))
            if phi0s is not None and perm_off[-1] > 0:
                phi0 = np.concatenate([np.asarray(phi0s[i]).reshape(-1, config.max_em_init, 3) for i, P, _, _ in grp if P > 0])
            elif on_device or perm_off[-1] == 0:
                phi0 = None
            else:
                from .simulations import gen_phi0s
                phi0 = gen_phi0s(rng or np.random.default_rng(), 2 * int(perm_off[-1]), config.max_em_init)
            if on_device or (phi0 is None and perm_off[-1] > 0):
                self.lib.set_seed(0 if seed is None else seed, self.device)
            cpg_off, rho_n, d_n, sub_off, sub_p, sub_q = self.pack_cpgs(refs)
            phi1 = np.array([items[i][0].phihat for i in idx]); phi2 = np.array([items[i][1].phihat for i in idx])
            res = self.lib.two_samp_test_batch(grp_off, Nl, rho_l, d_l, read_off, lu, lm, n1, phi1, phi2, perm_off, perm_ids, exact, phi0,
                                               config.max_em_init, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, config.max_em_iters,
                                               (AMAX, BMAX, CMAX), config.mstep_mode, want_null=want_null, device=self.device)
            t0 = 0
            for a, (i, P, ex, pr) in enumerate(grp):
                K = int(sub_off[a + 1] - sub_off[a])
                sl = slice(3 * sub_off[a], 3 * sub_off[a + 1])
                o = dict(T=res["T"][sl].reshape(3, K), pval=res["pval"][sl].reshape(3, K), cnt=None, n_valid=None, exact=None, coords=refs[a].nls_rgs)
                if P > 0:
                    o.update(cnt=res["cnt"][sl].reshape(3, K), n_valid=res["n_valid"][sl].reshape(3, K), exact=bool(ex))
                    if want_null:
                        o["T_null"] = res["T_null"][t0:t0 + P * 3 * K].reshape(P, 3, K)
                        o["fit_status"] = res["fit_status"][2 * perm_off[a]:2 * perm_off[a + 1]].reshape(P, 2)
                t0 += P * 3 * K
                outs[i] = o
        return outs


_default = None


def default_engine():
    global _default
    if _default is None:
        _default = Engine(0)
    return _default


# ---- single-struct entry points with the reference's names -----------------------------------------------
def get_phihat(rs, config, phi0s=None, rng=None, engine=None):
    (engine or default_engine()).get_phihat_batch([rs], config, None if phi0s is None else np.asarray(phi0s)[None], rng)


def rscle_grp_mod(rs):
    """src/nanopore.jl:269-280 — redefine groups as singletons (per-CpG chain).  The group-scale features are remembered in
    rs.group_view (the reference recomputes them with get_grp_info! when it goes "back to group scale",
    src/two_samp_perm_tests.jl:198-199)."""
    if rs.group_view is None and rs.L != rs.N:
        import copy
        gv = copy.copy(rs)
        gv.Nl, gv.rho_l, gv.dl = np.array(rs.Nl), np.array(rs.rho_l), np.array(rs.dl)
        rs.group_view = gv
    rs.L = rs.N
    rs.dl = np.asarray(rs.dn, dtype=np.float64)
    rs.rho_l = np.asarray(rs.rho_n, dtype=np.float64)
    rs.Nl = np.ones(rs.N)


def get_nls_reg_info(rs, config, engine=None):
    (engine or default_engine()).get_nls_reg_info(rs, config)


def get_stat_sums(rs, config, engine=None):
    (engine or default_engine()).get_stat_sums_batch([rs], config)


def comp_cmd(rs1, rs2, engine=None):
    return (engine or default_engine()).comp_cmd_pairs([rs1, rs2], [(0, 1)])[0]


def unmat_est_reg_test(ms_g1, ms_g2, engine=None, **kw):
    return (engine or default_engine()).unmat_est_reg_test_batch([ms_g1], [ms_g2], **kw)[0]


def mat_est_reg_test(ms_g1, ms_g2, engine=None, **kw):
    return (engine or default_engine()).mat_est_reg_test_batch([ms_g1], [ms_g2], **kw)[0]


def _copy_features(rs_ref):
    """cp_rs_flds (src/two_samp_perm_tests.jl:11-30): keep the genomic features, reset the data and results."""
    rs = RegStruct(chr=rs_ref.chr, chrst=rs_ref.chrst, chrend=rs_ref.chrend, N=rs_ref.N, L=rs_ref.L, Nl=np.array(rs_ref.Nl),
                   rho_n=np.array(rs_ref.rho_n), rho_l=np.array(rs_ref.rho_l), dn=np.array(rs_ref.dn), dl=np.array(rs_ref.dl),
                   cpg_pos=np.array(rs_ref.cpg_pos))
    return rs


def pmap_two_samp_null_stats(rs_ref, calls_s1, calls_s2, perm_ids, config, phi0s=None, rng=None, engine=None):
    """src/two_samp_perm_tests.jl:41-87 for ONE permutation: pooled reads are split by `perm_ids` (ascending
    1-based indices into vcat(calls_s1, calls_s2)), both halves are refitted, rescaled, summarised, compared.
    calls_* are (lu, lm) tuples of [m, L] arrays; phi0s [2, max_em_init, 3] replaces the two gen_ϕ0s draws."""
    eng = engine or default_engine()
    lu = np.concatenate([calls_s1[0], calls_s2[0]]); lm = np.concatenate([calls_s1[1], calls_s2[1]])
    idx = np.asarray(perm_ids, dtype=np.int64) - 1
    rest = np.setdiff1d(np.arange(lu.shape[0]), idx)            # deleteat!(aux_calls, perm_ids): the rest, in order
    a1, a2 = _copy_features(rs_ref), _copy_features(rs_ref)
    a1.set_calls(lu[idx], lm[idx]); a2.set_calls(lu[rest], lm[rest])
    K = rs_ref.nls_rgs.num
    nan3 = (np.full(K, np.nan), np.full(K, np.nan), np.full(K, np.nan))
    eng.get_phihat_batch([a1, a2], config, phi0s, rng)
    if not (len(a1.phihat) > 0 and len(a2.phihat) > 0):
        return nan3
    rscle_grp_mod(a1); rscle_grp_mod(a2)
    eng.get_stat_sums_batch([a1, a2], config)
    return a1.mml - a2.mml, a1.nme - a2.nme, eng.comp_cmd_pairs([a1, a2], [(0, 1)])[0]


def analyze_regions_wgbs(regs, calls, config, phi0s=None, rng=None, engine=None):
    """The body of pmap_analyze_reg_bam (src/wgbs.jl:335-389) for a batch of regions: `regs` carry the per-CpG features of
    get_grp_info!(rs, fa_rec, 1) (every CpG its own group), `calls[i]` is the int8 [M_i, N_i] table of the region's reads
    (−1 unmethylated, +1 methylated, 0 not covered — what get_calls_bam extracts from the XM tags: capi.BamHandle.calls,
    csrc/cpn_bam.cpp).  Filters of :363-367, then get_ϕhat_wgbs! and get_stat_sums! in one cpn_analyze_wgbs_batch call.  `engine`
    may be a callable returning the engine: it is only called when a region passes the filters."""
    keep = []
    for i, (rs, x) in enumerate(zip(regs, calls)):
        x = np.asarray(x, dtype=np.int8)
        N = len(rs.cpg_pos)
        rs.m = x.shape[0]
        if not (N > 9 and rs.m > 0 and x.shape[1] == N and N > 1):
            continue
        depth = np.count_nonzero(x) / N                       # get_ave_depth: observed cells / L
        frac = np.count_nonzero(np.any(x != 0, axis=0)) / N   # perc_gprs_obs: observed groups / L
        if depth >= config.min_cov and frac >= 0.66:
            keep.append(i)
    if not keep:
        return regs
    eng = engine() if callable(engine) else (engine or default_engine())
    sel = [regs[i] for i in keep]
    cpg_off = np.concatenate([[0], np.cumsum([len(rs.cpg_pos) for rs in sel])]).astype(np.int64)
    read_off = np.concatenate([[0], np.cumsum([np.asarray(calls[i]).shape[0] for i in keep])]).astype(np.int64)
    flat = np.concatenate([np.asarray(calls[i], dtype=np.int8).reshape(-1) for i in keep])
    rho_n = np.concatenate([rs.rho_n for rs in sel]); d_n = np.concatenate([rs.dn for rs in sel])
    for rs in sel:
        eng.get_nls_reg_info(rs, config)
    sub_off = np.concatenate([[0], np.cumsum([rs.nls_rgs.num for rs in sel])]).astype(np.int64)
    sub_p = np.concatenate([rs.nls_rgs.cpg_p for rs in sel]); sub_q = np.concatenate([rs.nls_rgs.cpg_q for rs in sel])
    n_init = config.max_em_init
    if phi0s is None and rng is not None:
        from . import simulations
        phi0s = simulations.gen_phi0s(rng, len(sel), n_init)
    elif phi0s is not None:
        phi0s = np.asarray(phi0s)[keep]
    out = eng.lib.analyze_wgbs_batch(cpg_off, rho_n, d_n, read_off, flat, phi0s, n_init, sub_off, sub_p, sub_q, config.max_em_iters,
                                     mstep_mode=config.mstep_mode, device=eng.device)
    for k, rs in enumerate(sel):
        if out["pick"][k] < 0:
            continue
        a, b, k0, k1 = cpg_off[k], cpg_off[k + 1], sub_off[k], sub_off[k + 1]
        rs.phihat = out["phi_hat"][k].copy(); rs.Q = float(out["Q"][k])
        rs.logZ = float(out["logZ"][k]); rs.ex = out["ex"][a:b].copy(); rs.exx = out["exx"][a - k:b - k - 1].copy()
        rs.log_gs = out["log_g"][k0:k1].copy(); rs.log_g1 = out["e_log_g1"][k0:k1].copy(); rs.log_g2 = out["e_log_g2"][k0:k1].copy()
        rs.mml = out["mml"][k0:k1].copy(); rs.nme = out["nme"][k0:k1].copy()
        rs.proc = True
    return regs


def analyze_regions(regs, config, phi0s=None, rng=None, engine=None):
    """The body of pmap_analyze_reg (src/nanopore.jl:292-352) for a whole chromosome's worth of regions at once —
    the batch point that replaces `pmap` at src/nanopore.jl:406.  `regs` already carry genomic features and calls
    (get_grp_info!/get_calls! stay on the Julia side).  Regions failing the reference's filters are left with proc=false."""
    eng = engine or default_engine()
    keep = [i for i, rs in enumerate(regs)
            if len(rs.cpg_pos) > 9 and rs.m > 0 and rs.get_ave_depth() >= config.min_cov and rs.perc_gprs_obs() >= 0.66 and rs.L > 1]
    if not keep:
        return regs
    sel = [regs[i] for i in keep]
    eng.get_phihat_batch(sel, config, None if phi0s is None else np.asarray(phi0s)[keep], rng)
    fitted = [rs for rs in sel if len(rs.phihat) > 0]
    for rs in fitted:
        rscle_grp_mod(rs)
    eng.get_stat_sums_batch(fitted, config)
    for rs in fitted:
        rs.proc = True
    return regs
