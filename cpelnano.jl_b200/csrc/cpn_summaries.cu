// cpn_summaries.cu — batched 2x2 transfer-matrix forward–backward pass and the MML / NME / CMD summaries.
//
// Reference path: get_stat_sums!  src/statistical_summaries.jl:470-500
//   log Z               get_rs_logZ! :75 → get_log_Z               src/transfer_matrix.jl:1199
//   E[X_n], E[X_nX_n+1] get_rs_exps! → get_E_X_log / get_E_XX_log  src/transfer_matrix.jl:1273 / :1322
//   log g1(±), log g2(±) comp_log_g_p / comp_log_g_q               src/statistical_summaries.jl:94 / :133
//   E[log g(X)]         get_rs_exp_log_g1! / g2!                   :212 / :284
//   MML, NME            comp_mml! :355, comp_nme! :388
//   CMD                 comp_cmd :512 with comp_crs_ent :427
//
// B200 design: one warp per fitted model.  Lanes own contiguous site segments; each lane multiplies the
// 2x2 transfer matrices of its segment, a warp-shuffle inclusive scan (prefix and suffix) gives every
// lane the message entering its segment from the left and from the right, and the lane then walks its
// segment once in each direction.  All products are kept in linear space with entries relative to the
// largest potential and renormalised by exact powers of two (integer exponents are carried separately),
// so the O(N) pass is overflow-free and replaces the reference's O(N²) log-domain re-multiplication.
// Every array a region needs is staged in shared memory; global reads/writes are lane-contiguous.
#include <cmath>
#include <cstdlib>

#include "cpn_common.cuh"

namespace {

constexpr int S_WARPS = 4;
constexpr int S_ARR = 11;                      // doubles of shared memory per site
constexpr double LN2 = 0.6931471805599453;
constexpr int CMD_MAXK = 32;                  // most sub-regions per region the lane-per-pair CMD kernel keeps in local memory

struct Mat2e { double m00, m01, m10, m11; int e; };   // value = m · 2^e, all entries ≥ 0

__device__ __forceinline__ Mat2e mat_identity() { return Mat2e{1.0, 0.0, 0.0, 1.0, 0}; }
__device__ __forceinline__ Mat2e mat_mul(const Mat2e& A, const Mat2e& B) {
    Mat2e C;
    C.m00 = fma(A.m00, B.m00, A.m01 * B.m10);
    C.m01 = fma(A.m00, B.m01, A.m01 * B.m11);
    C.m10 = fma(A.m10, B.m00, A.m11 * B.m10);
    C.m11 = fma(A.m10, B.m01, A.m11 * B.m11);
    int e;
    double sc = cpn_inv_pow2_of((C.m00 + C.m01) + (C.m10 + C.m11), &e);
    C.m00 *= sc; C.m01 *= sc; C.m10 *= sc; C.m11 *= sc;
    C.e = A.e + B.e + e;
    return C;
}
__device__ __forceinline__ Mat2e mat_shfl_up(const Mat2e& A, int o) {
    Mat2e B;
    B.m00 = __shfl_up_sync(0xffffffffu, A.m00, o); B.m01 = __shfl_up_sync(0xffffffffu, A.m01, o);
    B.m10 = __shfl_up_sync(0xffffffffu, A.m10, o); B.m11 = __shfl_up_sync(0xffffffffu, A.m11, o);
    B.e = __shfl_up_sync(0xffffffffu, A.e, o);
    return B;
}
__device__ __forceinline__ Mat2e mat_shfl_down(const Mat2e& A, int o) {
    Mat2e B;
    B.m00 = __shfl_down_sync(0xffffffffu, A.m00, o); B.m01 = __shfl_down_sync(0xffffffffu, A.m01, o);
    B.m10 = __shfl_down_sync(0xffffffffu, A.m10, o); B.m11 = __shfl_down_sync(0xffffffffu, A.m11, o);
    B.e = __shfl_down_sync(0xffffffffu, A.e, o);
    return B;
}

// Shared-memory view of one chain (per warp).
struct ChainSmem {
    double *wm, *wp, *t, *A0, *A1, *B0, *B1, *LSa, *LSb, *ex, *exx;
    __device__ ChainSmem(double* base, int stride) {
        wm = base; wp = base + stride; t = base + 2 * stride; A0 = base + 3 * stride; A1 = base + 4 * stride;
        B0 = base + 5 * stride; B1 = base + 6 * stride; LSa = base + 7 * stride; LSb = base + 8 * stride;
        ex = base + 9 * stride; exx = base + 10 * stride;
    }
};

// Builds, for the per-CpG chain with α_n = a + bρ_n, β_n = c/d_n (nanopore.jl:269-280):
//   A_n(x)  = Σ_{x_1..x_{n-1}} (everything left of n, incl. the coupling into n)   [= g1 at p=n]
//   B_n(x)  = Σ_{x_{n+1}..x_N} (everything right of n, incl. the coupling out of n) [= g2 at q=n]
// as A_n = Arel·exp(LSa_n), B_n = Brel·exp(LSb_n); returns log Z.  Also fills ex/exx.
__device__ double build_chain(ChainSmem& cs, int N, int lane, const double* __restrict__ rho, const double* __restrict__ dn, double a,
                              double b, double c) {
    const bool bpos = c >= 0.0;
    // (i) potentials, lanes strided over sites
    for (int n = lane; n < N; n += 32) {
        double alpha = a + b * rho[n];
        double e = exp(-2.0 * fabs(alpha));
        bool pos = alpha >= 0.0;
        cs.wm[n] = pos ? e : 1.0;
        cs.wp[n] = pos ? 1.0 : e;
        cs.LSa[n] = fabs(alpha);                                   // offsets, turned into prefix sums below
        double beta = (n > 0) ? c / dn[n - 1] : 0.0;
        cs.t[n] = exp(-2.0 * fabs(beta));
        cs.LSb[n] = fabs(beta);
    }
    __syncwarp();
    // (ii) per-lane segment product of T_n = Ψ_{n-1→n}·diag(φ_n)  (T_0 = diag(φ_0))
    const int seg = (N + 31) >> 5;
    const int s0 = min(N, lane * seg), s1 = min(N, s0 + seg);
    Mat2e P = mat_identity();
    double off_a = 0.0;                                            // Σ over the segment of |α_n| + |β into n|
    for (int n = s0; n < s1; ++n) {
        double t = cs.t[n], wm = cs.wm[n], wp = cs.wp[n];
        double same = bpos ? 1.0 : t, diff = bpos ? t : 1.0;
        Mat2e T;
        if (n == 0) T = Mat2e{wm, 0.0, 0.0, wp, 0};
        else T = Mat2e{same * wm, diff * wp, diff * wm, same * wp, 0};
        P = mat_mul(P, T);
        off_a += cs.LSa[n] + cs.LSb[n];
    }
    // (iii) inclusive prefix / suffix scans over lanes
    Mat2e pre = P, suf = P;
    double off_pre = off_a;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Mat2e up = mat_shfl_up(pre, o);
        double ou = __shfl_up_sync(0xffffffffu, off_pre, o);
        if (lane >= o) { pre = mat_mul(up, pre); off_pre += ou; }
        Mat2e dw = mat_shfl_down(suf, o);
        if (lane + o < 32) suf = mat_mul(suf, dw);
    }
    double off_tot = __shfl_sync(0xffffffffu, off_pre, 31);
    // exclusive versions
    Mat2e pre_ex = mat_shfl_up(pre, 1), suf_ex = mat_shfl_down(suf, 1);
    double off_ex = __shfl_up_sync(0xffffffffu, off_pre, 1);
    if (lane == 0) { pre_ex = mat_identity(); off_ex = 0.0; }
    if (lane == 31) suf_ex = mat_identity();
    // (iv) forward walk: F = 1ᵀ·T_0⋯T_n
    {
        double f0 = pre_ex.m00 + pre_ex.m10, f1 = pre_ex.m01 + pre_ex.m11;   // 1ᵀ·pre_ex
        int e = pre_ex.e;
        double off = off_ex;
        for (int n = s0; n < s1; ++n) {
            double t = cs.t[n];
            double same = bpos ? 1.0 : t, diff = bpos ? t : 1.0;
            double a0, a1;
            if (n == 0) { a0 = 1.0; a1 = 1.0; }
            else { a0 = same * f0 + diff * f1; a1 = diff * f0 + same * f1; }
            double oa = cs.LSa[n], ob = cs.LSb[n];
            off += ob;                                              // coupling into n belongs to A_n
            cs.A0[n] = a0; cs.A1[n] = a1;
            double lsa = (double)e * LN2 + off;
            off += oa;
            f0 = a0 * cs.wm[n]; f1 = a1 * cs.wp[n];
            int de;
            double sc = cpn_inv_pow2_of(f0 + f1, &de);
            f0 *= sc; f1 *= sc; e += de;
            cs.LSa[n] = lsa;                                        // overwrite the offset with the log-scale of A_n
            cs.ex[n] = oa;                                          // keep |α_n| for the backward walk
        }
    }
    __syncwarp();
    // backward walk: B_n = T_{n+1}⋯T_{N-1}·1
    {
        double b0 = suf_ex.m00 + suf_ex.m01, b1 = suf_ex.m10 + suf_ex.m11;   // suf_ex·1
        int e = suf_ex.e;
        // offsets to the right of this segment: total − inclusive prefix
        double off = off_tot - off_pre;
        for (int n = s1 - 1; n >= s0; --n) {
            cs.B0[n] = b0; cs.B1[n] = b1;
            double ob = cs.LSb[n];                                  // |β into n|
            cs.LSb[n] = (double)e * LN2 + off;
            // B_{n-1} = Ψ_{n-1→n}·(φ_n ∘ B_n)
            double c0 = cs.wm[n] * b0, c1 = cs.wp[n] * b1;
            double t = cs.t[n];
            double same = bpos ? 1.0 : t, diff = bpos ? t : 1.0;
            b0 = same * c0 + diff * c1; b1 = diff * c0 + same * c1;
            int de;
            double sc = cpn_inv_pow2_of(b0 + b1, &de);
            b0 *= sc; b1 *= sc; e += de;
            off += cs.ex[n] + ob;
        }
    }
    __syncwarp();
    // log Z from the last site: Z = Σ_x A_N(x) φ_N(x)
    double zl = cs.A0[N - 1] * cs.wm[N - 1] + cs.A1[N - 1] * cs.wp[N - 1];
    double absalpha_last = fabs(a + b * rho[N - 1]);
    double logZ = log(zl) + cs.LSa[N - 1] + absalpha_last;
    __syncwarp();
    // (v) marginals, lanes strided over sites
    for (int n = lane; n < N; n += 32) {
        double fm = cs.A0[n] * cs.wm[n], fp = cs.A1[n] * cs.wp[n];
        double pm = fm * cs.B0[n], pp = fp * cs.B1[n];
        cs.ex[n] = (pp - pm) / (pp + pm);
        if (n + 1 < N) {
            double cm = cs.wm[n + 1] * cs.B0[n + 1], cp = cs.wp[n + 1] * cs.B1[n + 1];
            double t = cs.t[n + 1];
            double same = bpos ? 1.0 : t, diff = bpos ? t : 1.0;
            double sm = same * (fm * cm + fp * cp), op = diff * (fm * cp + fp * cm);
            cs.exx[n] = (sm - op) / (sm + op);
        }
    }
    __syncwarp();
    return logZ;
}

__device__ __forceinline__ void log_g_at(const ChainSmem& cs, int N, int p, int q, double lg[4]) {
    // comp_log_g_p(p,∓1), comp_log_g_q(q,∓1): statistical_summaries.jl:94-161 (0 at the chain ends)
    if (p == 1) { lg[0] = 0.0; lg[1] = 0.0; }
    else { lg[0] = log(cs.A0[p - 1]) + cs.LSa[p - 1]; lg[1] = log(cs.A1[p - 1]) + cs.LSa[p - 1]; }
    if (q == N) { lg[2] = 0.0; lg[3] = 0.0; }
    else { lg[2] = log(cs.B0[q - 1]) + cs.LSb[q - 1]; lg[3] = log(cs.B1[q - 1]) + cs.LSb[q - 1]; }
}

// What comp_crs_ent (statistical_summaries.jl:427-459) needs from model i for sub-region [p, q] (1-based): the sums
// Σ E_i[X_n], Σ ρ_n E_i[X_n], Σ E_i[X_n X_n+1]/d_n — so that α̃·E_i[X] + β̃·E_i[XX] = ã·S0 + b̃·S1 + c̃·S2 for any mixture
// parameters ϕ̃ — and the boundary marginals E_i[X_p], E_i[X_q].  Pair-independent: computed once per model, shared by every
// pair the model is in (11 in a 6-vs-6 comparison).
__device__ __forceinline__ void model_sums(const double* ex, const double* exx, const double* __restrict__ rho, const double* __restrict__ dn,
                                           int p, int q, double ms[5]) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int n = p; n <= q; ++n) { const double e = ex[n - 1]; s0 += e; s1 = fma(rho[n - 1], e, s1); }
    for (int n = p; n <= q - 1; ++n) s2 += exx[n - 1] / dn[n - 1];
    ms[0] = s0; ms[1] = s1; ms[2] = s2; ms[3] = ex[p - 1]; ms[4] = ex[q - 1];
}

__global__ void __launch_bounds__(S_WARPS * 32) k_stat_sums(int64_t T, int stride, const double* __restrict__ phi,
                                                             const int64_t* __restrict__ reg_of, const int64_t* __restrict__ ex_off,
                                                             const int64_t* __restrict__ k_off,
                                                             const int64_t* __restrict__ cpg_off, const double* __restrict__ rho_n,
                                                             const double* __restrict__ d_n, const int64_t* __restrict__ sub_off,
                                                             const int64_t* __restrict__ sub_p, const int64_t* __restrict__ sub_q,
                                                             double* __restrict__ logZ_out, double* __restrict__ ex_out,
                                                             double* __restrict__ exx_out, double* __restrict__ log_g_out,
                                                             double* __restrict__ elg1_out, double* __restrict__ elg2_out,
                                                             double* __restrict__ mml_out, double* __restrict__ nme_out,
                                                             double* __restrict__ msum_out /*[k_off[T]][5] or null*/) {
    extern __shared__ double smem[];
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t tm = blockIdx.x * (int64_t)S_WARPS + w;
    if (tm >= T) return;
    ChainSmem cs(smem + (size_t)w * S_ARR * stride, stride);
    int64_t r = reg_of ? reg_of[tm] : tm;
    int64_t n0 = cpg_off[r];
    int N = (int)(cpg_off[r + 1] - n0);
    const double* rho = rho_n + n0;
    const double* dn = d_n + (n0 - r);
    double a = phi[3 * tm], b = phi[3 * tm + 1], c = phi[3 * tm + 2];
    double logZ = build_chain(cs, N, lane, rho, dn, a, b, c);
    if (lane == 0) logZ_out[tm] = logZ;
    if (ex_out) {                                        // the fused comparison paths keep E[X], E[XX] on chip (msum is all they need)
        int64_t e0 = ex_off[tm];
        for (int n = lane; n < N; n += 32) {
            ex_out[e0 + n] = cs.ex[n];
            if (n + 1 < N) exx_out[e0 - tm + n] = cs.exx[n];
        }
    }
    int64_t ks = sub_off[r];
    int K = (int)(sub_off[r + 1] - ks);
    int64_t k0 = k_off[tm];
    for (int k = lane; k < K; k += 32) {
        int p = (int)sub_p[ks + k], q = (int)sub_q[ks + k];
        double lg[4] = {cpn_nan(), cpn_nan(), cpn_nan(), cpn_nan()};
        double e1 = cpn_nan(), e2 = cpn_nan(), mml = cpn_nan(), nme = cpn_nan();
        double ms[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        if (p != 0 && q != 0) {
            if (msum_out) model_sums(cs.ex, cs.exx, rho, dn, p, q, ms);
            log_g_at(cs, N, p, q, lg);
            double pp_p = 0.5 * (1.0 + cs.ex[p - 1]), pp_q = 0.5 * (1.0 + cs.ex[q - 1]);
            e1 = (p == 1) ? 0.0 : fma(pp_p, lg[1], (1.0 - pp_p) * lg[0]);        // E[log g1(X_p)]  (:212-236)
            e2 = (q == N) ? 0.0 : fma(pp_q, lg[3], (1.0 - pp_q) * lg[2]);        // E[log g2(X_q)]  (:284-308)
            double Nk = (double)(q - p + 1);
            double s = 0.0, dot1 = 0.0, dot2 = 0.0;
            for (int n = p; n <= q; ++n) {
                double exn = cs.ex[n - 1];
                s += exn;
                dot1 = fma(a + b * rho[n - 1], exn, dot1);
            }
            for (int n = p; n <= q - 1; ++n) dot2 = fma(c / dn[n - 1], cs.exx[n - 1], dot2);
            mml = 0.5 / Nk * (Nk + s);                                               // comp_mml! :369
            nme = (logZ - e1 - e2 - dot1 - dot2) / (Nk * LN2);                       // comp_nme! :403-405
        } else if (q != 0) {
            e2 = (q == N) ? 0.0 : cpn_nan();
        }
        log_g_out[4 * (k0 + k) + 0] = lg[0]; log_g_out[4 * (k0 + k) + 1] = lg[1];
        log_g_out[4 * (k0 + k) + 2] = lg[2]; log_g_out[4 * (k0 + k) + 3] = lg[3];
        elg1_out[k0 + k] = e1; elg2_out[k0 + k] = e2; mml_out[k0 + k] = mml; nme_out[k0 + k] = nme;
        if (msum_out) {
#pragma unroll
            for (int j = 0; j < 5; ++j) msum_out[5 * (k0 + k) + j] = ms[j];
        }
    }
}

// comp_cmd for one pair per warp: chain of the geometric mixture ϕ̃ = (ϕa+ϕb)/2 (:517-533), cross
// entropies against the two models' stored E[X], E[XX] (:427-459), CMD (:543-565).
__global__ void __launch_bounds__(S_WARPS * 32) k_cmd(int64_t P, int stride, const int64_t* __restrict__ pair_a,
                                                       const int64_t* __restrict__ pair_b, const double* __restrict__ phi_tbl,
                                                       const int64_t* __restrict__ reg_of, const int64_t* __restrict__ ex_off,
                                                       const double* __restrict__ ex_tbl, const double* __restrict__ exx_tbl,
                                                       const int64_t* __restrict__ nme_off, const double* __restrict__ nme_tbl,
                                                       const int64_t* __restrict__ cpg_off, const double* __restrict__ rho_n,
                                                       const double* __restrict__ d_n, const int64_t* __restrict__ sub_off,
                                                       const int64_t* __restrict__ sub_p, const int64_t* __restrict__ sub_q,
                                                       const int64_t* __restrict__ cmd_off, double* __restrict__ cmd) {
    extern __shared__ double smem[];
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t pi = blockIdx.x * (int64_t)S_WARPS + w;
    if (pi >= P) return;
    ChainSmem cs(smem + (size_t)w * S_ARR * stride, stride);
    int64_t ma = pair_a[pi], mb = pair_b[pi];
    int64_t r = reg_of[ma];
    int64_t n0 = cpg_off[r];
    int N = (int)(cpg_off[r + 1] - n0);
    const double* rho = rho_n + n0;
    const double* dn = d_n + (n0 - r);
    double a = 0.5 * (phi_tbl[3 * ma] + phi_tbl[3 * mb]), b = 0.5 * (phi_tbl[3 * ma + 1] + phi_tbl[3 * mb + 1]),
           c = 0.5 * (phi_tbl[3 * ma + 2] + phi_tbl[3 * mb + 2]);
    double logZm = build_chain(cs, N, lane, rho, dn, a, b, c);
    const double* exa = ex_tbl + ex_off[ma];
    const double* exb = ex_tbl + ex_off[mb];
    const double* exxa = exx_tbl + (ex_off[ma] - ma);
    const double* exxb = exx_tbl + (ex_off[mb] - mb);
    const double* nmea = nme_tbl + nme_off[ma];
    const double* nmeb = nme_tbl + nme_off[mb];
    int64_t k0 = sub_off[r];
    int K = (int)(sub_off[r + 1] - k0);
    double* out = cmd + cmd_off[pi];
    for (int k = lane; k < K; k += 32) {
        int p = (int)sub_p[k0 + k], q = (int)sub_q[k0 + k];
        double res = cpn_nan();
        if (p != 0 && q != 0) {
            double lg[4];
            log_g_at(cs, N, p, q, lg);
            double ce[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double* exi = i ? exb : exa;
                const double* exxi = i ? exxb : exxa;
                double pp_p = 0.5 * (1.0 + exi[p - 1]), pp_q = 0.5 * (1.0 + exi[q - 1]);
                double e1 = (p == 1) ? 0.0 : fma(pp_p, lg[1], (1.0 - pp_p) * lg[0]);
                double e2 = (q == N) ? 0.0 : fma(pp_q, lg[3], (1.0 - pp_q) * lg[2]);
                double dot1 = 0.0, dot2 = 0.0;
                for (int n = p; n <= q; ++n) dot1 = fma(a + b * rho[n - 1], exi[n - 1], dot1);
                for (int n = p; n <= q - 1; ++n) dot2 = fma(c / dn[n - 1], exxi[n - 1], dot2);
                ce[i] = logZm - dot1 - e1 - e2 - dot2;                              // comp_crs_ent :448-449
            }
            double len = (double)(q - p + 1);
            double ent1 = nmea[k] * len * LN2, ent2 = nmeb[k] * len * LN2;          // :554-555
            if (!(ce[0] == 0.0 && ce[1] == 0.0)) res = 1.0 - (ent1 + ent2) / (ce[0] + ce[1]);   // :560-563
        }
        out[k] = res;
    }
}


// model_sums for models whose E[X], E[XX] tables came from the caller (cpn_cmd_batch): thread per (model, sub-region).
__global__ void k_msum_from_ex(int64_t T, const int64_t* __restrict__ reg_of, const int64_t* __restrict__ ex_off, const double* __restrict__ ex_tbl,
                               const double* __restrict__ exx_tbl, const int64_t* __restrict__ k_off, const int64_t* __restrict__ cpg_off,
                               const double* __restrict__ rho_n, const double* __restrict__ d_n, const int64_t* __restrict__ sub_off,
                               const int64_t* __restrict__ sub_p, const int64_t* __restrict__ sub_q, double* __restrict__ msum) {
    int64_t t = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (t >= T) return;
    const int64_t r = reg_of[t], n0 = cpg_off[r], ks = sub_off[r];
    const int K = (int)(sub_off[r + 1] - ks);
    for (int k = lane; k < K; k += 32) {
        const int p = (int)sub_p[ks + k], q = (int)sub_q[ks + k];
        double ms[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        if (p != 0 && q != 0) model_sums(ex_tbl + ex_off[t], exx_tbl + (ex_off[t] - t), rho_n + n0, d_n + (n0 - r), p, q, ms);
#pragma unroll
        for (int j = 0; j < 5; ++j) msum[5 * (k_off[t] + k) + j] = ms[j];
    }
}

__global__ void k_check_ordered(int64_t R, const int64_t* __restrict__ sub_off, const int64_t* __restrict__ sub_p, const int64_t* __restrict__ sub_q,
                                const int64_t* __restrict__ cpg_off, int* __restrict__ bad) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= R) return;
    bool ok = sub_off[r + 1] - sub_off[r] <= CMD_MAXK;
    int64_t last = 0;
    const int64_t N = cpg_off[r + 1] - cpg_off[r];
    for (int64_t k = sub_off[r]; k < sub_off[r + 1]; ++k) {
        const int64_t p = sub_p[k], q = sub_q[k];
        if (p == 0 || q == 0) continue;
        if (p <= last || q < p || q > N) ok = false;
        last = q;
    }
    if (!ok) *bad = 1;
}

// (ρ_n, 1/d_{n-1}) per CpG site, 16 bytes: what a chain step of k_cmd_lane loads (one broadcast LDG.128 per warp and site).
__global__ void k_region_tables(int64_t R, const int64_t* __restrict__ cpg_off, const double* __restrict__ rho_n, const double* __restrict__ d_n,
                                double2* __restrict__ rd) {
    int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (r >= R) return;
    const int64_t n0 = cpg_off[r];
    const int N = (int)(cpg_off[r + 1] - n0);
    for (int n = lane; n < N; n += 32) rd[n0 + n] = make_double2(rho_n[n0 + n], n > 0 ? 1.0 / d_n[n0 - r + n - 1] : 0.0);
}

// comp_cmd (statistical_summaries.jl:512-570), ONE LANE PER PAIR.  The warp-per-pair kernel above spends most of its time in
// shuffle scans and shared-memory staging for chains of ≈ 60 sites; here a lane walks the mixture chain (ϕ̃ = (ϕa+ϕb)/2, :517-527)
// once, left to right, carrying the forward message F and the 2×2 product G of the transfer matrices since the last
// sub-region end.  That is all comp_crs_ent (:427-459) needs when the sub-regions are ordered and disjoint, as
// get_nls_reg_info! makes them:  log g̃1(x_p) comes from F just before site p; log g̃2(x_q) = log[(G_{k+1}⋯G_last·1)(x_q)] from a
// backward pass over the ≤ K stored products; α̃·E_i[X] + β̃·E_i[XX] = ã·S0 + b̃·S1 + c̃·S2 with the per-model sums of model_sums.
// Messages are linear-space doubles relative to the larger potential, rescaled by exact powers of two.
__global__ void __launch_bounds__(128) k_cmd_lane(int64_t P, const int64_t* __restrict__ pair_a, const int64_t* __restrict__ pair_b,
                                                  const double* __restrict__ phi_tbl, const int64_t* __restrict__ reg_of,
                                                  const int64_t* __restrict__ k_off, const double* __restrict__ msum,
                                                  const double* __restrict__ nme_tbl, const int64_t* __restrict__ cpg_off,
                                                  const double2* __restrict__ rd, const int64_t* __restrict__ sub_off,
                                                  const int64_t* __restrict__ sub_p, const int64_t* __restrict__ sub_q,
                                                  const int64_t* __restrict__ cmd_off, double* __restrict__ cmd) {
    const int64_t pi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (pi >= P) return;
    const int64_t ma = pair_a[pi], mb = pair_b[pi], r = reg_of[ma];
    const int64_t n0 = cpg_off[r];
    const int N = (int)(cpg_off[r + 1] - n0);
    const double2* rdr = rd + n0;
    const double a = 0.5 * (phi_tbl[3 * ma] + phi_tbl[3 * mb]), b = 0.5 * (phi_tbl[3 * ma + 1] + phi_tbl[3 * mb + 1]),
                 c = 0.5 * (phi_tbl[3 * ma + 2] + phi_tbl[3 * mb + 2]);
    const bool bpos = c >= 0.0;
    const double cabs2 = -2.0 * fabs(c);
    const int64_t ks = sub_off[r];
    const int K = (int)(sub_off[r + 1] - ks);
    // per sub-region: log g̃1(∓) at p, the product G since the previous sub-region end (4 entries, exponent, offset)
    double lg1m[CMD_MAXK], lg1p[CMD_MAXK], g00[CMD_MAXK], g01[CMD_MAXK], g10[CMD_MAXK], g11[CMD_MAXK], goff[CMD_MAXK];
    int gexp[CMD_MAXK];
    int k = 0;                                            // next non-empty sub-region whose p or q lies ahead
    while (k < K && (sub_p[ks + k] == 0 || sub_q[ks + k] == 0)) ++k;
    int pk = k < K ? (int)sub_p[ks + k] : 0, qk = k < K ? (int)sub_q[ks + k] : 0;
    double f0 = 1.0, f1 = 1.0, offF = 0.0;               // forward message through site n-1 (incl. φ), value = f·2^eF·e^offF
    int eF = 0;
    double G00 = 1.0, G01 = 0.0, G10 = 0.0, G11 = 1.0, offG = 0.0;
    int eG = 0;
    for (int n = 0; n < N; ++n) {
        const double2 x = __ldg(rdr + n);
        const double alpha = fma(b, x.x, a);
        const double absa = fabs(alpha);
        const double e = cpn_exp_neg(-2.0 * absa);
        const bool pos = alpha >= 0.0;
        const double wm = pos ? e : 1.0, wp = pos ? 1.0 : e;
        double a0, a1, absb = 0.0;
        if (n == 0) { a0 = 1.0; a1 = 1.0; }
        else {
            const double bt = cabs2 * x.y;                  // -2|β|, β = c/d_{n-1}
            absb = -0.5 * bt;
            const double t = cpn_exp_neg(bt);
            const double same = bpos ? 1.0 : t, diff = bpos ? t : 1.0;
            a0 = fma(diff, f1, same * f0); a1 = fma(diff, f0, same * f1);
            // G ← G·Ψ
            const double h00 = fma(G01, diff, G00 * same), h01 = fma(G01, same, G00 * diff);
            const double h10 = fma(G11, diff, G10 * same), h11 = fma(G11, same, G10 * diff);
            G00 = h00; G01 = h01; G10 = h10; G11 = h11;
        }
        offF += absb;                                       // the coupling into n belongs to A_n (everything left of n)
        if (n + 1 == pk) {                                  // comp_log_g_p(p, ∓1) :94-122; 0 at the chain start
            const double ls = (double)eF * LN2 + offF;
            lg1m[k] = (n == 0) ? 0.0 : log(a0) + ls;
            lg1p[k] = (n == 0) ? 0.0 : log(a1) + ls;
        }
        offF += absa; offG += absb + absa;
        f0 = a0 * wm; f1 = a1 * wp;
        if (n == 0) { G00 = wm; G11 = wp; }                 // T_0 = diag(φ_0)
        else { G00 *= wm; G10 *= wm; G01 *= wp; G11 *= wp; }
        if (n & 1) {                                        // rescale every second site (entries ≤ 1 per step, ≥ e^-2|·| ... safe)
            int de;
            double sc = cpn_inv_pow2_of(f0 + f1, &de);
            f0 *= sc; f1 *= sc; eF += de;
            sc = cpn_inv_pow2_of((G00 + G01) + (G10 + G11), &de);
            G00 *= sc; G01 *= sc; G10 *= sc; G11 *= sc; eG += de;
        }
        if (n + 1 == qk) {                                  // close the segment that ends at q_k
            g00[k] = G00; g01[k] = G01; g10[k] = G10; g11[k] = G11; gexp[k] = eG; goff[k] = offG;
            G00 = 1.0; G01 = 0.0; G10 = 0.0; G11 = 1.0; eG = 0; offG = 0.0;
            ++k;
            while (k < K && (sub_p[ks + k] == 0 || sub_q[ks + k] == 0)) ++k;
            pk = k < K ? (int)sub_p[ks + k] : 0; qk = k < K ? (int)sub_q[ks + k] : 0;
        }
    }
    const double logZm = log(f0 + f1) + (double)eF * LN2 + offF;
    // backward over the sub-regions: B = (product of everything right of q_k)·1, starting from the tail after the last q
    double b0 = G00 + G01, b1 = G10 + G11, offB = offG;
    int eB = eG;
    const double* msa = msum + 5 * k_off[ma];
    const double* msb = msum + 5 * k_off[mb];
    const double* nmea = nme_tbl + k_off[ma];
    const double* nmeb = nme_tbl + k_off[mb];
    double* out = cmd + cmd_off[pi];
    for (int kk = K - 1; kk >= 0; --kk) {
        const int p = (int)sub_p[ks + kk], q = (int)sub_q[ks + kk];
        if (p == 0 || q == 0) { out[kk] = cpn_nan(); continue; }
        const double lsb = (double)eB * LN2 + offB;
        const double lg2m = (q == N) ? 0.0 : log(b0) + lsb, lg2p = (q == N) ? 0.0 : log(b1) + lsb;   // comp_log_g_q :133-161
        double ce[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const double* ms = (i ? msb : msa) + 5 * kk;
            const double pp_p = 0.5 * (1.0 + ms[3]), pp_q = 0.5 * (1.0 + ms[4]);
            const double e1 = (p == 1) ? 0.0 : fma(pp_p, lg1p[kk], (1.0 - pp_p) * lg1m[kk]);       // E_i[log g̃1(X_p)] :248-275
            const double e2 = (q == N) ? 0.0 : fma(pp_q, lg2p, (1.0 - pp_q) * lg2m);             // E_i[log g̃2(X_q)] :320-347
            const double dot1 = fma(b, ms[1], a * ms[0]), dot2 = c * ms[2];
            ce[i] = logZm - dot1 - e1 - e2 - dot2;                                                  // comp_crs_ent :448-449
        }
        const double len = (double)(q - p + 1);
        const double ent1 = nmea[kk] * len * LN2, ent2 = nmeb[kk] * len * LN2;                    // :554-555
        out[kk] = (ce[0] == 0.0 && ce[1] == 0.0) ? cpn_nan() : 1.0 - (ent1 + ent2) / (ce[0] + ce[1]);   // :560-563
        // B ← G_kk·B: everything right of q_{kk-1}
        const double c0 = fma(g01[kk], b1, g00[kk] * b0), c1 = fma(g11[kk], b1, g10[kk] * b0);
        int de;
        const double sc = cpn_inv_pow2_of(c0 + c1, &de);
        b0 = c0 * sc; b1 = c1 * sc; eB += gexp[kk] + de; offB += goff[kk];
    }
}

inline unsigned blocks_for(int64_t n, int per_block) { return (unsigned)((n + per_block - 1) / per_block); }

}  // namespace

int cpn_stat_sums_launch(cpn_ctx* ctx, int64_t T, const double* phi, const int64_t* reg_of, const int64_t* ex_off, const int64_t* k_off,
                         const int64_t* cpg_off, const double* rho_n, const double* d_n, const int64_t* sub_off, const int64_t* sub_p,
                         const int64_t* sub_q, int64_t Nmax, double* logZ, double* ex, double* exx, double* log_g, double* e_log_g1,
                         double* e_log_g2, double* mml, double* nme, double* msum) {
    int stride = (int)((Nmax + 1) | 1);                              // odd stride: fewer shared-memory bank conflicts
    size_t smem = (size_t)S_WARPS * S_ARR * stride * sizeof(double);
    CPN_ARG(ctx, smem <= 200 * 1024, "region with too many CpG sites for the shared-memory staging (N > ~500)");
    CPN_CUDA(ctx, cudaFuncSetAttribute(k_stat_sums, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        KTimer kt(ctx, CPN_K_STATSUM);
        k_stat_sums<<<blocks_for(T, S_WARPS), S_WARPS * 32, smem, ctx->stream>>>(T, stride, phi, reg_of, ex_off, k_off, cpg_off, rho_n, d_n, sub_off,
                                                                                  sub_p, sub_q, logZ, ex, exx, log_g, e_log_g1, e_log_g2, mml, nme, msum);
    }
    CPN_CUDA(ctx, cudaGetLastError());
    return CPN_OK;
}

bool cpn_subregions_ordered(int64_t R, const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q, const int64_t* cpg_off) {
    for (int64_t r = 0; r < R; ++r) {
        if (sub_off[r + 1] - sub_off[r] > CMD_MAXK) return false;
        int64_t last = 0;
        const int64_t N = cpg_off[r + 1] - cpg_off[r];
        for (int64_t k = sub_off[r]; k < sub_off[r + 1]; ++k) {
            const int64_t p = sub_p[k], q = sub_q[k];
            if (p == 0 || q == 0) continue;
            if (p <= last || q < p || q > N) return false;
            last = q;
        }
    }
    return true;
}

// The same check with the tables on the device (one flag comes back).
int cpn_subregions_ordered_device(cpn_ctx* ctx, int64_t R, const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q, const int64_t* cpg_off,
                                  bool* ordered) {
    DevBuf<int> flag;
    CPN_CUDA(ctx, flag.alloc(1, ctx->stream));
    CPN_CUDA(ctx, cudaMemsetAsync(flag.p, 0, sizeof(int), ctx->stream));
    { KTimer kt(ctx, CPN_K_PREP); k_check_ordered<<<blocks_for(R, 256), 256, 0, ctx->stream>>>(R, sub_off, sub_p, sub_q, cpg_off, flag.p); }
    int h = 1;
    CPN_CUDA(ctx, cudaMemcpyAsync(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CPN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *ordered = h == 0;
    return CPN_OK;
}

int cpn_cmd_lane_launch(cpn_ctx* ctx, int64_t P, const int64_t* pair_a, const int64_t* pair_b, const double* phi_tbl, const int64_t* reg_of,
                        const int64_t* k_off, const double* msum, const double* nme_tbl, int64_t R, const int64_t* cpg_off, const double* rho_n,
                        const double* d_n, double2* rd_scratch /*[sumN]*/, const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q,
                        const int64_t* cmd_off, double* cmd) {
    { KTimer kt(ctx, CPN_K_PREP);
      k_region_tables<<<blocks_for(R * 32, 256), 256, 0, ctx->stream>>>(R, cpg_off, rho_n, d_n, rd_scratch); }
    { KTimer kt(ctx, CPN_K_CMD);
      k_cmd_lane<<<blocks_for(P, 128), 128, 0, ctx->stream>>>(P, pair_a, pair_b, phi_tbl, reg_of, k_off, msum, nme_tbl, cpg_off, rd_scratch, sub_off,
                                                              sub_p, sub_q, cmd_off, cmd); }
    CPN_CUDA(ctx, cudaGetLastError());
    return CPN_OK;
}

// get_stat_sums! of T models followed by comp_cmd of P pairs of them, everything on the device.  With ordered sub-regions (the
// get_nls_reg_info! case) the lane-per-pair kernel runs and E[X], E[XX] are never written to memory; otherwise the general
// warp-per-pair kernel reads them back from a scratch table.
int cpn_summaries_and_cmd(cpn_ctx* ctx, bool ordered, int64_t T, const double* phi, const int64_t* reg_of, const int64_t* ex_off,
                          const int64_t* k_off, int64_t sN /*ex_off[T]*/, int64_t sK /*k_off[T]*/, int64_t R, int64_t sumN, const int64_t* cpg_off,
                          const double* rho_n, const double* d_n, const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q, int64_t Nmax,
                          int64_t P, const int64_t* pair_a, const int64_t* pair_b, const int64_t* cmd_off, double* logZ, double* log_g,
                          double* e_log_g1, double* e_log_g2, double* mml, double* nme, double* cmd) {
    cudaStream_t s = ctx->stream;
    if (ordered) {
        DevBuf<double> msum;
        DevBuf<double2> rd;
        CPN_CUDA(ctx, msum.alloc(5 * sK, s));
        CPN_CUDA(ctx, rd.alloc(sumN, s));
        int rc = cpn_stat_sums_launch(ctx, T, phi, reg_of, ex_off, k_off, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, Nmax, logZ, nullptr, nullptr, log_g,
                                      e_log_g1, e_log_g2, mml, nme, msum.p);
        if (rc) return rc;
        return cpn_cmd_lane_launch(ctx, P, pair_a, pair_b, phi, reg_of, k_off, msum.p, nme, R, cpg_off, rho_n, d_n, rd.p, sub_off, sub_p, sub_q, cmd_off,
                                   cmd);
    }
    DevBuf<double> ex, exx;
    CPN_CUDA(ctx, ex.alloc(sN, s));
    CPN_CUDA(ctx, exx.alloc(sN - T, s));
    int rc = cpn_stat_sums_launch(ctx, T, phi, reg_of, ex_off, k_off, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, Nmax, logZ, ex.p, exx.p, log_g, e_log_g1,
                                  e_log_g2, mml, nme, nullptr);
    if (rc) return rc;
    return cpn_cmd_launch(ctx, P, pair_a, pair_b, phi, reg_of, ex_off, ex.p, exx.p, k_off, nme, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, Nmax, cmd_off,
                          cmd);
}

int cpn_cmd_launch(cpn_ctx* ctx, int64_t P, const int64_t* pair_a, const int64_t* pair_b, const double* phi_tbl, const int64_t* reg_of,
                   const int64_t* ex_off, const double* ex_tbl, const double* exx_tbl, const int64_t* nme_off, const double* nme_tbl,
                   const int64_t* cpg_off, const double* rho_n, const double* d_n, const int64_t* sub_off, const int64_t* sub_p,
                   const int64_t* sub_q, int64_t Nmax, const int64_t* cmd_off, double* cmd) {
    int stride = (int)((Nmax + 1) | 1);
    size_t smem = (size_t)S_WARPS * S_ARR * stride * sizeof(double);
    CPN_ARG(ctx, smem <= 200 * 1024, "region with too many CpG sites for the shared-memory staging (N > ~500)");
    CPN_CUDA(ctx, cudaFuncSetAttribute(k_cmd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        KTimer kt(ctx, CPN_K_CMD);
        k_cmd<<<blocks_for(P, S_WARPS), S_WARPS * 32, smem, ctx->stream>>>(P, stride, pair_a, pair_b, phi_tbl, reg_of, ex_off, ex_tbl, exx_tbl,
                                                                            nme_off, nme_tbl, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, cmd_off,
                                                                            cmd);
    }
    CPN_CUDA(ctx, cudaGetLastError());
    return CPN_OK;
}

extern "C" {

int cpn_stat_sums_batch(cpn_ctx* ctx, int64_t T, const double* phi, const int64_t* reg_of, const int64_t* ex_off, const int64_t* k_off,
                        int64_t R, const int64_t* cpg_off, const double* rho_n, const double* d_n, const int64_t* sub_off,
                        const int64_t* sub_p, const int64_t* sub_q, double* logZ, double* ex, double* exx, double* log_g,
                        double* e_log_g1, double* e_log_g2, double* mml, double* nme) {
    if (!ctx) return CPN_ERR_ARG;
    CPN_ARG(ctx, T >= 0 && R >= 0 && phi && cpg_off && rho_n && d_n && sub_off && sub_p && sub_q, "null input pointer");
    CPN_ARG(ctx, logZ && ex && exx && log_g && e_log_g1 && e_log_g2 && mml && nme, "null output pointer");
    CPN_ARG(ctx, reg_of ? (ex_off && k_off) : (T == R), "reg_of needs ex_off and k_off; without reg_of T must equal R");
    CPN_CUDA(ctx, cudaSetDevice(ctx->device));
    LaunchCounter lc(ctx);
    if (T == 0) return CPN_OK;
    if (!reg_of) { ex_off = cpg_off; k_off = sub_off; }
    int64_t sumN = cpg_off[R], sumK = sub_off[R], Nmax = 0;
    for (int64_t r = 0; r < R; ++r) {
        int64_t N = cpg_off[r + 1] - cpg_off[r];
        CPN_ARG(ctx, N >= 2, "every region needs N >= 2 CpG sites");
        Nmax = std::max(Nmax, N);
    }
    if (reg_of)
        for (int64_t t = 0; t < T; ++t) {
            CPN_ARG(ctx, reg_of[t] >= 0 && reg_of[t] < R, "reg_of out of range");
            CPN_ARG(ctx, ex_off[t + 1] - ex_off[t] == cpg_off[reg_of[t] + 1] - cpg_off[reg_of[t]], "ex_off inconsistent with the model's region");
            CPN_ARG(ctx, k_off[t + 1] - k_off[t] == sub_off[reg_of[t] + 1] - sub_off[reg_of[t]], "k_off inconsistent with the model's region");
        }
    const int64_t outN = ex_off[T], outK = k_off[T];
    cudaStream_t s = ctx->stream;
    CPN_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    {
        DevBuf<double> d_phi, d_rho, d_dn, o_logZ, o_ex, o_exx, o_lg, o_e1, o_e2, o_mml, o_nme;
        DevBuf<int64_t> d_cpg, d_so, d_sp, d_sq, d_reg, d_exo, d_ko;
        CPN_CUDA(ctx, d_phi.upload(phi, 3 * T, s)); CPN_CUDA(ctx, d_cpg.upload(cpg_off, R + 1, s));
        CPN_CUDA(ctx, d_rho.upload(rho_n, sumN, s)); CPN_CUDA(ctx, d_dn.upload(d_n, sumN - R, s));
        CPN_CUDA(ctx, d_so.upload(sub_off, R + 1, s)); CPN_CUDA(ctx, d_sp.upload(sub_p, sumK, s)); CPN_CUDA(ctx, d_sq.upload(sub_q, sumK, s));
        if (reg_of) {
            CPN_CUDA(ctx, d_reg.upload(reg_of, T, s)); CPN_CUDA(ctx, d_exo.upload(ex_off, T + 1, s)); CPN_CUDA(ctx, d_ko.upload(k_off, T + 1, s));
        }
        CPN_CUDA(ctx, o_logZ.alloc(T, s)); CPN_CUDA(ctx, o_ex.alloc(outN, s)); CPN_CUDA(ctx, o_exx.alloc(outN - T, s));
        CPN_CUDA(ctx, o_lg.alloc(4 * outK, s)); CPN_CUDA(ctx, o_e1.alloc(outK, s)); CPN_CUDA(ctx, o_e2.alloc(outK, s));
        CPN_CUDA(ctx, o_mml.alloc(outK, s)); CPN_CUDA(ctx, o_nme.alloc(outK, s));
        int rc = cpn_stat_sums_launch(ctx, T, d_phi.p, reg_of ? d_reg.p : nullptr, reg_of ? d_exo.p : d_cpg.p, reg_of ? d_ko.p : d_so.p, d_cpg.p,
                                      d_rho.p, d_dn.p, d_so.p, d_sp.p, d_sq.p, Nmax, o_logZ.p, o_ex.p, o_exx.p, o_lg.p, o_e1.p, o_e2.p, o_mml.p,
                                      o_nme.p, nullptr);
        if (rc) return rc;
        CPN_CUDA(ctx, o_logZ.download(logZ, T, s)); CPN_CUDA(ctx, o_ex.download(ex, outN, s)); CPN_CUDA(ctx, o_exx.download(exx, outN - T, s));
        CPN_CUDA(ctx, o_lg.download(log_g, 4 * outK, s)); CPN_CUDA(ctx, o_e1.download(e_log_g1, outK, s));
        CPN_CUDA(ctx, o_e2.download(e_log_g2, outK, s)); CPN_CUDA(ctx, o_mml.download(mml, outK, s)); CPN_CUDA(ctx, o_nme.download(nme, outK, s));
        CPN_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
        CPN_CUDA(ctx, cudaStreamSynchronize(s));
    }
    float ms = 0;
    CPN_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    if (ctx->profile) cpn_profile_collect(ctx);
    return CPN_OK;
}

int cpn_cmd_batch(cpn_ctx* ctx, int64_t P, const int64_t* pair_a, const int64_t* pair_b, int64_t T, const double* phi_tbl,
                  const int64_t* reg_of, const int64_t* ex_off, const double* ex_tbl, const double* exx_tbl, const int64_t* nme_off,
                  const double* nme_tbl, int64_t R, const int64_t* cpg_off, const double* rho_n, const double* d_n, const int64_t* sub_off,
                  const int64_t* sub_p, const int64_t* sub_q, const int64_t* cmd_off, double* cmd) {
    if (!ctx) return CPN_ERR_ARG;
    CPN_ARG(ctx, P >= 0 && T >= 0 && R >= 0, "negative size");
    CPN_ARG(ctx, pair_a && pair_b && phi_tbl && reg_of && ex_off && ex_tbl && exx_tbl && nme_off && nme_tbl, "null model table pointer");
    CPN_ARG(ctx, cpg_off && rho_n && d_n && sub_off && sub_p && sub_q && cmd_off && cmd, "null pointer");
    CPN_CUDA(ctx, cudaSetDevice(ctx->device));
    LaunchCounter lc(ctx);
    if (P == 0) return CPN_OK;
    for (int64_t i = 0; i < P; ++i) {
        CPN_ARG(ctx, pair_a[i] >= 0 && pair_a[i] < T && pair_b[i] >= 0 && pair_b[i] < T, "pair index out of range");
        CPN_ARG(ctx, reg_of[pair_a[i]] == reg_of[pair_b[i]], "both models of a pair must belong to the same region");
    }
    int64_t sumN = cpg_off[R], sumK = sub_off[R], Nmax = 0;
    for (int64_t r = 0; r < R; ++r) Nmax = std::max(Nmax, cpg_off[r + 1] - cpg_off[r]);
    int stride = (int)((Nmax + 1) | 1);
    size_t smem = (size_t)S_WARPS * S_ARR * stride * sizeof(double);
    CPN_ARG(ctx, smem <= 200 * 1024, "region with too many CpG sites for the shared-memory staging (N > ~500)");
    cudaStream_t s = ctx->stream;
    CPN_CUDA(ctx, cudaFuncSetAttribute(k_cmd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CPN_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    {
        DevBuf<int64_t> d_pa, d_pb, d_reg, d_exo, d_nmo, d_cpg, d_so, d_sp, d_sq, d_co;
        DevBuf<double> d_phi, d_ex, d_exx, d_nme, d_rho, d_dn, o_cmd;
        CPN_CUDA(ctx, d_pa.upload(pair_a, P, s)); CPN_CUDA(ctx, d_pb.upload(pair_b, P, s)); CPN_CUDA(ctx, d_reg.upload(reg_of, T, s));
        CPN_CUDA(ctx, d_exo.upload(ex_off, T + 1, s)); CPN_CUDA(ctx, d_nmo.upload(nme_off, T + 1, s));
        CPN_CUDA(ctx, d_phi.upload(phi_tbl, 3 * T, s)); CPN_CUDA(ctx, d_ex.upload(ex_tbl, ex_off[T], s));
        CPN_CUDA(ctx, d_exx.upload(exx_tbl, ex_off[T] - T, s)); CPN_CUDA(ctx, d_nme.upload(nme_tbl, nme_off[T], s));
        CPN_CUDA(ctx, d_cpg.upload(cpg_off, R + 1, s)); CPN_CUDA(ctx, d_rho.upload(rho_n, sumN, s)); CPN_CUDA(ctx, d_dn.upload(d_n, sumN - R, s));
        CPN_CUDA(ctx, d_so.upload(sub_off, R + 1, s)); CPN_CUDA(ctx, d_sp.upload(sub_p, sumK, s)); CPN_CUDA(ctx, d_sq.upload(sub_q, sumK, s));
        CPN_CUDA(ctx, d_co.upload(cmd_off, P + 1, s)); CPN_CUDA(ctx, o_cmd.alloc(cmd_off[P], s));
        const char* e_lane = getenv("CPN_CMD_LANE");
        if (!(e_lane && atoi(e_lane) == 0) && cpn_subregions_ordered(R, sub_off, sub_p, sub_q, cpg_off)) {
            // per-model sums from the caller's E[X], E[XX] tables (the arithmetic cpn_stat_sums_launch uses when it produces them
            // itself, so the staged and the fused routes agree bit for bit), then one lane per pair
            DevBuf<double> msum;
            DevBuf<double2> rd;
            CPN_CUDA(ctx, msum.alloc(5 * nme_off[T], s)); CPN_CUDA(ctx, rd.alloc(sumN, s));
            { KTimer kt(ctx, CPN_K_PREP);
              k_msum_from_ex<<<blocks_for(T * 32, 256), 256, 0, s>>>(T, d_reg.p, d_exo.p, d_ex.p, d_exx.p, d_nmo.p, d_cpg.p, d_rho.p, d_dn.p, d_so.p,
                                                                     d_sp.p, d_sq.p, msum.p); }
            int rc = cpn_cmd_lane_launch(ctx, P, d_pa.p, d_pb.p, d_phi.p, d_reg.p, d_nmo.p, msum.p, d_nme.p, R, d_cpg.p, d_rho.p, d_dn.p, rd.p, d_so.p,
                                         d_sp.p, d_sq.p, d_co.p, o_cmd.p);
            if (rc) return rc;
        } else {
            KTimer kt(ctx, CPN_K_CMD);
            k_cmd<<<blocks_for(P, S_WARPS), S_WARPS * 32, smem, s>>>(P, stride, d_pa.p, d_pb.p, d_phi.p, d_reg.p, d_exo.p, d_ex.p, d_exx.p, d_nmo.p,
                                                                      d_nme.p, d_cpg.p, d_rho.p, d_dn.p, d_so.p, d_sp.p, d_sq.p, d_co.p, o_cmd.p);
        }
        CPN_CUDA(ctx, cudaGetLastError());
        CPN_CUDA(ctx, o_cmd.download(cmd, cmd_off[P], s));
        CPN_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
        CPN_CUDA(ctx, cudaStreamSynchronize(s));
    }
    float ms = 0;
    CPN_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    if (ctx->profile) cpn_profile_collect(ctx);
    return CPN_OK;
}

}  // extern "C"
