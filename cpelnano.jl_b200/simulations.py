"""Synthetic nanopore workloads — a vectorised restatement of the reference's simulator.

Host-side workload generator used by tests/ and bench.py (never on the product's compute path):
  gen_x_mc        src/simulations.jl:512-573   exact sequential sampler of the Ising chain
  cpel_samp_ont   src/simulations.jl:674-714   x̄ → Gaussian current y → (log p(y|x=-1), log p(y|x=+1))
  gen_ϕ0s         src/em_algorithm.jl:65-78    random EM starts on the 0.01 grid
  gen_grp_comp_data src/simulations.jl:588-652 per-sample ϕ = ϕ_group + noise
Levels μ_u=80, μ_m=84, σ=2 follow calls/PaperSimulations/Aim-4/Aim-4-Null-Pvals.jl:24-29.
The reference draws from Julia's global MersenneTwister; here every array is a deterministic function
of (seed, shapes) through numpy's PCG64, so oracle and GPU runs see identical inputs.
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
MASTER_SEED = 20211111


def load_layouts(path=None):
    """Real 3 kb CpG layouts (N>9, L>1) of the bundled hg38 chr17 window: package data (data/layouts_targeted.npz, written by
    tests/golden/make_fixtures.py from the reference's targeted-example FASTA)."""
    path = path or os.path.join(_HERE, "data", "layouts_targeted.npz")
    z = np.load(path)
    return {k: z[k] for k in z.files}


def gen_phi0s(rng, R, n_init):
    """em_algorithm.jl:72 — a∈-5:0.01:5, b∈-50:0.01:50, c∈0:0.01:10; integer-then-divide gives the decimal grid."""
    a = (rng.integers(0, 1001, size=(R, n_init)) - 500) / 100.0
    b = (rng.integers(0, 10001, size=(R, n_init)) - 5000) / 100.0
    c = rng.integers(0, 1001, size=(R, n_init)) / 100.0
    return np.stack([a, b, c], axis=-1).astype(np.float64)


def sample_chain(rng, alpha, beta, lens, M):
    """Exact samples of x ∈ {-1,+1}^L for a batch of chains.

    alpha [R, Lmax], beta [R, Lmax-1] (zero padded), lens [R]; returns int8 [R, M, Lmax].
    Backward messages b_l(x) = Σ_{x'} exp(β_l x x' + α_{l+1} x') b_{l+1}(x'), kept as the log-ratio
    h_l = ½ log(b_l(+)/b_l(-)) = atanh(tanh β_l · tanh(α_{l+1} + h_{l+1})); then
    P(x_l=+1 | x_{l-1}) = σ(2(α_l + β_{l-1} x_{l-1} + h_l))  — the same law as gen_x_mc.
    """
    R, Lmax = alpha.shape
    h = np.zeros((R, Lmax))
    for l in range(Lmax - 2, -1, -1):
        live = (l + 1) < lens
        t = np.tanh(beta[:, l]) * np.tanh(alpha[:, l + 1] + h[:, l + 1])
        h[:, l] = np.where(live, np.arctanh(np.clip(t, -1 + 1e-16, 1 - 1e-16)), 0.0)
    x = np.zeros((R, M, Lmax), dtype=np.int8)
    prev = np.zeros((R, M))
    for l in range(Lmax):
        field = (alpha[:, l] + h[:, l])[:, None] + (beta[:, l - 1][:, None] * prev if l > 0 else 0.0)
        p = 1.0 / (1.0 + np.exp(-2.0 * field))
        xl = np.where(rng.random((R, M)) < p, 1, -1)
        x[:, :, l] = xl
        prev = xl.astype(np.float64)
    return x


def _pad(values, off, Lmax, fill=0.0):
    R = len(off) - 1
    out = np.full((R, Lmax), fill)
    lens = np.diff(off)
    idx = np.arange(Lmax)[None, :] < lens[:, None]
    out[idx] = values
    return out


class NanoBatch:
    """CSR batch of estimation regions at CpG-group scale plus the per-CpG features — the packed
    inputs of cpn_fit_batch / cpn_stat_sums_batch (include/cpelnano_b200.h)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def R(self):
        return len(self.grp_off) - 1

    def region(self, r):
        g0, g1 = self.grp_off[r], self.grp_off[r + 1]
        L = g1 - g0
        M = self.read_off[r + 1] - self.read_off[r]
        o0 = self.obs_off[r]
        return dict(L=int(L), M=int(M), Nl=self.Nl[g0:g1], rho_l=self.rho_l[g0:g1], d_l=self.d_l[g0 - r:g1 - r - 1],
                    lu=self.log_pyx_u[o0:o0 + M * L].reshape(M, L), lm=self.log_pyx_m[o0:o0 + M * L].reshape(M, L))

    def subset(self, idx):
        idx = np.asarray(idx)
        keys_g, keys_n = ("Nl", "rho_l"), ("rho_n", "cpg_pos")
        out = {}
        gl = [np.arange(self.grp_off[r], self.grp_off[r + 1]) for r in idx]
        dl = [np.arange(self.grp_off[r] - r, self.grp_off[r + 1] - r - 1) for r in idx]
        nl = [np.arange(self.cpg_off[r], self.cpg_off[r + 1]) for r in idx]
        dn = [np.arange(self.cpg_off[r] - r, self.cpg_off[r + 1] - r - 1) for r in idx]
        ol = [np.arange(self.obs_off[r], self.obs_off[r + 1]) for r in idx]
        for k in keys_g:
            out[k] = getattr(self, k)[np.concatenate(gl)]
        out["d_l"] = self.d_l[np.concatenate(dl)]
        for k in keys_n:
            out[k] = getattr(self, k)[np.concatenate(nl)]
        out["d_n"] = self.d_n[np.concatenate(dn)]
        out["log_pyx_u"] = self.log_pyx_u[np.concatenate(ol)]
        out["log_pyx_m"] = self.log_pyx_m[np.concatenate(ol)]
        out["grp_off"] = np.concatenate([[0], np.cumsum([len(g) for g in gl])]).astype(np.int64)
        out["cpg_off"] = np.concatenate([[0], np.cumsum([len(g) for g in nl])]).astype(np.int64)
        out["obs_off"] = np.concatenate([[0], np.cumsum([len(g) for g in ol])]).astype(np.int64)
        Ms = np.diff(self.read_off)[idx]
        out["read_off"] = np.concatenate([[0], np.cumsum(Ms)]).astype(np.int64)
        out["chrst"] = self.chrst[idx]; out["chrend"] = self.chrend[idx]
        out["phi_true"] = self.phi_true[idx]
        return NanoBatch(**out)


def make_nano_batch(R, M=30, seed=MASTER_SEED, pobs=0.9, mu_m=84.0, sigma_m=2.0, mu_u=80.0, sigma_u=2.0, layouts=None,
                    phi_true=None, chunk=4096):
    """cycle through the real layouts; ϕ ~ U(-1,1)×U(-20,20)×U(0,5); M reads per region."""
    lay = layouts or load_layouts()
    nlay = len(lay["grp_off"]) - 1
    rng = np.random.default_rng(seed)
    which = np.arange(R) % nlay
    Ls = np.diff(lay["grp_off"])[which]
    Ns = np.diff(lay["cpg_off"])[which]
    grp_off = np.concatenate([[0], np.cumsum(Ls)]).astype(np.int64)
    cpg_off = np.concatenate([[0], np.cumsum(Ns)]).astype(np.int64)

    def gather(vals, off, shrink):
        parts = [vals[off[w] - (w if shrink else 0):off[w + 1] - (w if shrink else 0) - (1 if shrink else 0)] for w in range(nlay)]
        return np.concatenate([parts[w] for w in which])

    Nl = gather(lay["Nl"], lay["grp_off"], False); rho_l = gather(lay["rho_l"], lay["grp_off"], False)
    d_l = gather(lay["d_l"], lay["grp_off"], True)
    rho_n = gather(lay["rho_n"], lay["cpg_off"], False); d_n = gather(lay["d_n"], lay["cpg_off"], True)
    cpg_pos = gather(lay["cpg_pos"], lay["cpg_off"], False)
    if phi_true is None:
        phi_true = np.stack([rng.uniform(-1, 1, R), rng.uniform(-20, 20, R), rng.uniform(0, 5, R)], axis=1)
    read_off = (np.arange(R + 1) * M).astype(np.int64)
    obs_off = np.concatenate([[0], np.cumsum(Ls * M)]).astype(np.int64)
    lu = np.empty(obs_off[-1]); lm = np.empty(obs_off[-1])
    d_off = grp_off - np.arange(R + 1)
    c_u = -np.log(sigma_u) - 0.5 * np.log(2 * np.pi); c_m = -np.log(sigma_m) - 0.5 * np.log(2 * np.pi)
    for s in range(0, R, chunk):
        e = min(R, s + chunk)
        Lmax = int(Ls[s:e].max())
        off = grp_off[s:e + 1] - grp_off[s]
        Nl_p = _pad(Nl[grp_off[s]:grp_off[e]], off, Lmax); rho_p = _pad(rho_l[grp_off[s]:grp_off[e]], off, Lmax)
        doff = d_off[s:e + 1] - d_off[s]
        dl_p = _pad(d_l[d_off[s]:d_off[e]], doff, Lmax - 1, fill=1.0)
        ph = phi_true[s:e]
        alpha = Nl_p * (ph[:, 0:1] + ph[:, 1:2] * rho_p)
        beta = ph[:, 2:3] / dl_p
        lens = Ls[s:e]
        x = sample_chain(rng, alpha, beta, lens, M)                       # [r, M, Lmax]
        obs = rng.random(x.shape) < pobs
        y = np.where(x == 1, rng.normal(mu_m, sigma_m, x.shape), rng.normal(mu_u, sigma_u, x.shape))
        lpu = c_u - 0.5 * ((y - mu_u) / sigma_u) ** 2
        lpm = c_m - 0.5 * ((y - mu_m) / sigma_m) ** 2
        lpu[~obs] = np.nan; lpm[~obs] = np.nan
        valid = np.broadcast_to((np.arange(Lmax)[None, :] < lens[:, None])[:, None, :], x.shape)
        lu[obs_off[s]:obs_off[e]] = lpu[valid]
        lm[obs_off[s]:obs_off[e]] = lpm[valid]
    chrst = lay["chrst"][which]; chrend = lay["chrend"][which]
    return NanoBatch(grp_off=grp_off, cpg_off=cpg_off, read_off=read_off, obs_off=obs_off, Nl=Nl, rho_l=rho_l, d_l=d_l, rho_n=rho_n,
                     d_n=d_n, cpg_pos=cpg_pos, log_pyx_u=lu, log_pyx_m=lm, chrst=chrst, chrend=chrend, phi_true=phi_true)


def wgbs_obs_pattern(rng, M, L, pobs):
    """Which groups a WGBS read observes, as cpel_samp_wgbs draws it (simulations.jl:737-763): walking along the chain a new
    fragment starts with probability pobs whenever at least 15 groups have passed since the last start, and a fragment covers its
    first 9 groups (obs_count 1..9)."""
    obs = np.zeros((M, L), dtype=bool)
    for m in range(M):
        cnt = 15
        for l in range(L):
            if rng.random() < pobs and cnt >= 15:
                cnt = 0
            cnt += 1
            obs[m, l] = cnt < 10
    return obs


def make_wgbs_batch(R, M=40, seed=MASTER_SEED, pobs=0.25, layouts=None, phi_true=None):
    """WGBS analogue of make_nano_batch at CpG resolution (every CpG site its own group: N_l = 1, ρ_l = ρ_n, d_l = d_n, as
    pmap_analyze_reg_bam builds its regions, wgbs.jl:351): exact samples of the chain, hard calls x ∈ {−1, 0, +1} (int8, 0 = not
    covered by the read) with cpel_samp_wgbs's coverage pattern.  Returns a NanoBatch whose group arrays are the per-CpG ones, with
    `calls` [Σ M·N] and the equivalent emission tables log_pyx_u / log_pyx_m (0 / −50, NaN where not covered)."""
    lay = layouts or load_layouts()
    nlay = len(lay["cpg_off"]) - 1
    rng = np.random.default_rng(seed)
    which = np.arange(R) % nlay
    Ns = np.diff(lay["cpg_off"])[which]
    cpg_off = np.concatenate([[0], np.cumsum(Ns)]).astype(np.int64)
    rho_n = np.concatenate([lay["rho_n"][lay["cpg_off"][w]:lay["cpg_off"][w + 1]] for w in which])
    d_n = np.concatenate([lay["d_n"][lay["cpg_off"][w] - w:lay["cpg_off"][w + 1] - w - 1] for w in which])
    cpg_pos = np.concatenate([lay["cpg_pos"][lay["cpg_off"][w]:lay["cpg_off"][w + 1]] for w in which])
    if phi_true is None:
        phi_true = np.stack([rng.uniform(-1, 1, R), rng.uniform(-20, 20, R), rng.uniform(0, 5, R)], axis=1)
    read_off = (np.arange(R + 1) * M).astype(np.int64)
    obs_off = np.concatenate([[0], np.cumsum(Ns * M)]).astype(np.int64)
    calls = np.zeros(obs_off[-1], dtype=np.int8)
    for r in range(R):
        N = int(Ns[r])
        rho = rho_n[cpg_off[r]:cpg_off[r + 1]]
        dn = d_n[cpg_off[r] - r:cpg_off[r + 1] - r - 1]
        alpha = (phi_true[r, 0] + phi_true[r, 1] * rho)[None, :]
        beta = (phi_true[r, 2] / dn)[None, :]
        x = sample_chain(rng, alpha, beta, np.array([N]), M)[0]          # [M, N]
        obs = wgbs_obs_pattern(rng, M, N, pobs)
        calls[obs_off[r]:obs_off[r + 1]] = np.where(obs, x, 0).reshape(-1)
    lu = np.where(calls == 1, -50.0, np.where(calls == -1, 0.0, np.nan))
    lm = np.where(calls == 1, 0.0, np.where(calls == -1, -50.0, np.nan))
    return NanoBatch(grp_off=cpg_off, cpg_off=cpg_off, read_off=read_off, obs_off=obs_off, Nl=np.ones(int(cpg_off[-1])), rho_l=rho_n, d_l=d_n,
                     rho_n=rho_n, d_n=d_n, cpg_pos=cpg_pos, log_pyx_u=lu, log_pyx_m=lm, calls=calls, chrst=lay["chrst"][which],
                     chrend=lay["chrend"][which], phi_true=phi_true)


def gen_grp_phis(rng, phi1, phi2, s1, s2, nse_sc=1.0):
    """simulations.jl:608,631 — ϕ_s = ϕ_g + nse_sc·[U{-0.1:0.01:0.1}, U{-0.5:0.01:0.5}, U{0:0.01:0.2}]; phi1,phi2 [R,3]."""
    R = phi1.shape[0]

    def noise(S):
        return np.stack([(rng.integers(0, 21, size=(S, R)) - 10) / 100.0, (rng.integers(0, 101, size=(S, R)) - 50) / 100.0,
                         rng.integers(0, 21, size=(S, R)) / 100.0], axis=-1)

    return phi1[None] + nse_sc * noise(s1), phi2[None] + nse_sc * noise(s2)


def layouts_from_fasta(lib, chromosome, size_est_reg=3000, min_grp_dist=10, targets=None):
    """Layout tables (the `layouts=` argument of make_nano_batch) for every estimation region of a chromosome that passes the
    N > 9, L > 1 filters of pmap_analyze_reg, from the native feature extraction; also returns the regions' group intervals."""
    from . import capi
    part = chromosome.chr_part_native(lib, size_est_reg, min_grp_dist, targets)
    st = np.array([p[0] for p in part], dtype=np.int64); en = np.array([p[1] for p in part], dtype=np.int64)
    o = capi.grp_info_batch(lib, chromosome.cg, chromosome.size, st, en, min_grp_dist)
    N, L = np.diff(o["cpg_off"]), np.diff(o["grp_off"])
    keep = np.nonzero((N > 9) & (L > 1))[0]
    cat = lambda v, off: np.concatenate([v[off[r]:off[r + 1]] for r in keep])      # noqa: E731
    lay = dict(chrst=st[keep], chrend=en[keep], cpg_off=np.concatenate([[0], np.cumsum(N[keep])]).astype(np.int64),
               grp_off=np.concatenate([[0], np.cumsum(L[keep])]).astype(np.int64), cpg_pos=cat(o["cpg_pos"], o["cpg_off"]),
               rho_n=cat(o["rho_n"], o["cpg_off"]), d_n=cat(o["d_n"], o["dn_off"]), Nl=cat(o["Nl"], o["grp_off"]),
               rho_l=cat(o["rho_l"], o["grp_off"]), d_l=cat(o["d_l"], o["dl_off"]))
    return lay, cat(o["grp_st"], o["grp_off"]), cat(o["grp_end"], o["grp_off"])


def write_nanopolish_tsv(path, nb, chrom_name, grp_st, grp_end, fmt=repr):
    """A NanoBatch as a nanopolish call-methylation file: one line per observed (read, group), sorted by start; reads are
    region-local and named r<region>_<read>.  grp_st / grp_end are the group intervals of the batch's regions in order (group
    interval = [first CpG − 5, last CpG + 5], so the 0-based nanopolish start / end are grp_st + 4 / grp_end − 6).  fmt = repr
    writes doubles that parse back bit-identically; fmt = "{:.2f}".format gives nanopolish's two decimals."""
    with open(path, "w") as f:
        f.write("chromosome\tstrand\tstart\tend\tread_name\tlog_lik_ratio\tlog_lik_methylated\tlog_lik_unmethylated\tnum_calling_strands\t"
                "num_motifs\tsequence\n")
        for r in range(nb.R):
            g0, g1 = int(nb.grp_off[r]), int(nb.grp_off[r + 1])
            L, M = g1 - g0, int(nb.read_off[r + 1] - nb.read_off[r])
            lu = nb.log_pyx_u[nb.obs_off[r]:nb.obs_off[r + 1]].reshape(M, L); lm = nb.log_pyx_m[nb.obs_off[r]:nb.obs_off[r + 1]].reshape(M, L)
            for l in range(L):
                st, en = int(grp_st[g0 + l]) + 4, int(grp_end[g0 + l]) - 6
                for m in range(M):
                    if lu[m, l] == lu[m, l]:
                        f.write(f"{chrom_name}\t+\t{st}\t{en}\tr{r:06d}_{m:03d}\t{fmt(float(lm[m, l] - lu[m, l]))}\t{fmt(float(lm[m, l]))}\t"
                                f"{fmt(float(lu[m, l]))}\t1\t1\tNNNNNCGNNNNN\n")
