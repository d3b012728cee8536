// cpn_em.cu — multi-start EM for batches of estimation regions (sm_100a).
//
// Reference path: get_ϕhat! → em_inst → em_iter   src/em_algorithm.jl:371/304/132
//   E-step  get_cond_exs_log + get_ess             src/transfer_matrix.jl:944-965, src/em_algorithm.jl:99-113
//   M-step  NLsolve on U(ϕ) = ∇logZ(ϕ) − ES/M       src/em_algorithm.jl:152-178, src/transfer_matrix.jl:1225-1246
//   score   comp_ave_log_py → get_logpy_lgtrck      src/em_algorithm.jl:263-288, src/transfer_matrix.jl:843-870
//
// B200 design:
//   * O(M·L) instead of the reference's O(M·L²): E[S(X)|y_m] = ∇_ϕ log Zc_m(ϕ), evaluated by ONE forward
//     sweep of the 2-state chain that carries the 3 tangents (∂a,∂b,∂c) next to the message; no backward
//     pass, no message storage.  Messages are kept in linear (not log) space as doubles renormalised by
//     exact powers of two; potentials are expressed relative to their larger entry so nothing overflows.
//   * E-step: one warp per (region,start) EM instance, lanes = reads; emission likelihood ratios are
//     pre-packed once per call as [chunk][group][32 lanes] so every step is one coalesced 256 B load;
//     per-group potentials of the instance are staged in shared memory by the warp (lanes = groups).
//   * M-step: one lane per EM instance (the solve is a serial recursion along the chain); gradient AND
//     Jacobian of log Z come from one forward sweep carrying first- and second-order tangents, feeding a
//     restatement of NLsolve's trust-region dogleg.
//   * The EM loop is a sequence of batched launches over a compacted list of still-active instances.
#include <algorithm>
#include <array>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <chrono>
#include <cstdlib>
#include <thread>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>

#include "cpn_common.cuh"
#include "cpn_philox.h"

namespace {

constexpr int TILE = 64;           // groups staged per shared-memory tile
#ifndef E_WARPS_N
#define E_WARPS_N 4
#endif
constexpr int E_WARPS = E_WARPS_N; // warps (EM instances) per E-step block
constexpr int M_THREADS = 128;     // lanes (EM instances) per M-step block
#ifndef WS_BLOCKS
#define WS_BLOCKS 4                 // resident M-step blocks per SM the register budget is tuned for
#endif
#ifndef WS_ROUNDS
#define WS_ROUNDS 2                 // take-work + bookkeeping rounds per loop trip of the M-step
#endif
constexpr double LLR_CLAMP = 300.0;
constexpr double CBRT_EPS = 6.0554544523933395e-6;

// One 32-byte record per CpG group for the M-step sweep (ϕ-independent): N_l, N_lρ_l, 1/d and 1/d² of the coupling INTO
// the group (0 for the first group).  Small on purpose: the records of all chains resident on an SM must stay in L1.
struct __align__(32) SweepSite { double na, nb, id, id2; };

// What an E-step warp needs to know about its EM instance, stored by POSITION in the active list (written by
// k_init_state / k_compact_scatter) so that the warp's first loads do not depend on each other:
// instance id, region geometry, start of the packed emissions, and the instance's current ϕ.
struct __align__(32) ItemDesc { int inst, L, M, ridx; int64_t g0, ew_off; };
// A "region" of the EM kernels.  Normally one estimation region with its own reads.  A VIEW (two-sample permutation
// refits) shares the group tables and the packed emissions of a base region and names its reads through an index table:
// ridx ≥ 0 is the first row (32 int32 read ids, one per lane and chunk) of the view in RegionTables::ridx; −1 = all reads.
// llrmax = largest |log-likelihood ratio| among the region's emissions (after the clamp), idmax = largest 1/d_l: together with |c|
// they bound how fast an E-step message can grow or shrink per group, i.e. how many groups may pass between two rescalings.
struct __align__(16) RegionDesc { int64_t g0, ew_off, rb_off; int32_t L, M, ridx; float llrmax, idmax; int32_t pad; };
static_assert(sizeof(RegionDesc) == 48, "RegionDesc layout");
constexpr int EW_ROW = 32;         // doubles per emission row: one likelihood ratio per lane

struct RegionTables {
    // per group (offsets grp_off): N_l, N_l·ρ_l and 1/d_l (coupling l→l+1; entry L-1 unused)
    const int64_t* grp_off;
    const double *na, *nb, *invd;
    const SweepSite* sq;       // per group: what the second-order sweep of the M-step needs, one 32-byte record
    const RegionDesc* rdesc;   // [R]
    // emission likelihood ratios r = exp(log p(y|+1) − log p(y|−1)), 1 for unobserved / padding lanes
    const int64_t* read_off;   // [R+1]
    const int64_t* ew_off;     // [R] start of the region's block: [chunk][l][32]
    const int64_t* rb_off;     // [R] start of the per-read base (Σ_obs log p(y|−1)) block: [chunk][32]
    const double *ew, *rb;
    const double* ewi;         // 1/r, same layout as ew (the E-step reads it for blocks of groups whose heavier state is x=+1)
    const int32_t* ridx;       // read-id rows of views (null when there are none)
};

// Per-group potentials of one EM instance, staged in shared memory as three 16-byte pairs so that a chain step reads
// them with three LDS.128 (all lanes read the same address: broadcast).
struct SiteTile {
    double2 w[TILE];     // (φ(−1), φ(+1)) relative to max(φ)
    double2 tn[TILE];    // (t = ψ(opposite)/ψ(aligned) for |β| into this group, N_l)
    double2 bd[TILE];    // (N_l·ρ_l, ±1/d_l of the coupling into this group; sign = sign(c))
};

// ---------------------------------------------------------------------------------------------------
// prep kernels
// ---------------------------------------------------------------------------------------------------
// link_functions.jl:16-25 inputs: α_l = a·N_l + b·(N_l ρ_l), β_l = c·(1/d_l)
__global__ void k_prep_sites(int64_t R, const int64_t* __restrict__ grp_off, const double* __restrict__ Nl,
                             const double* __restrict__ rho, const double* __restrict__ dl, double* __restrict__ na,
                             double* __restrict__ nb, double* __restrict__ invd, SweepSite* __restrict__ sq) {
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= R) return;
    int64_t g0 = grp_off[warp], g1 = grp_off[warp + 1];
    int64_t L = g1 - g0;
    for (int64_t l = lane; l < L; l += 32) {
        double n = Nl[g0 + l];
        double b = n * rho[g0 + l];
        double id = (l < L - 1) ? 1.0 / dl[g0 - warp + l] : 0.0;
        na[g0 + l] = n;
        nb[g0 + l] = b;
        invd[g0 + l] = id;
        double idin = (l > 0) ? 1.0 / dl[g0 - warp + l - 1] : 0.0;
        SweepSite ss;
        ss.na = n; ss.nb = b; ss.id = idin; ss.id2 = idin * idin;
        sq[g0 + l] = ss;
    }
}

// Packs emissions (structures.jl:121-128: obs, log_pyx_u, log_pyx_m) as likelihood ratios r = exp(log p(y|+1) − log p(y|−1))
// and their reciprocals (two planes of the same layout), lanes = reads, and records per region the bounds
// RegionDesc::llrmax / idmax.
// |LLR| is clamped at 300 with the dropped mass (< e^-300 relative) moved into the per-read base term.
__global__ void k_prep_emis(int64_t R, const int64_t* __restrict__ grp_off, const int64_t* __restrict__ read_off,
                            const int64_t* __restrict__ obs_off, const double* __restrict__ lu, const double* __restrict__ lm, bool packed,
                            const int64_t* __restrict__ ew_off, const int64_t* __restrict__ rb_off, const double* __restrict__ invd,
                            double* __restrict__ ew, double* __restrict__ ewi, double* __restrict__ rb, RegionDesc* __restrict__ rdesc) {
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= R) return;
    const int64_t g0 = grp_off[warp];
    int64_t L = grp_off[warp + 1] - g0;
    int64_t M = read_off[warp + 1] - read_off[warp];
    int64_t nchunk = (M + 31) >> 5;
    const double* lu_r = lu + obs_off[warp];
    const double* lm_r = packed ? lm + read_off[warp] : lm + obs_off[warp];      // packed: one base per read
    double dmax = 0.0;
    for (int64_t ch = 0; ch < nchunk; ++ch) {
        int64_t m = ch * 32 + lane;
        double base = 0.0;
        double* dst = ew + ew_off[warp] + ch * L * 32 + lane;
        double* dsti = ewi + ew_off[warp] + ch * L * 32 + lane;
        for (int64_t l = 0; l < L; ++l) {
            double r = 1.0, ri = 1.0;
            if (m < M) {
                // packed: u is the log-likelihood ratio itself and the read's Σ log p(y|−1) arrives ready-made; only what the
                // clamp moves out of the ratio is added to it
                double u = lu_r[m * L + l], v = packed ? 0.0 : lm_r[m * L + l];
                if (u == u) {   // observed
                    double d = packed ? u : v - u;
                    if (packed) u = 0.0;
                    if (d > LLR_CLAMP) { u = packed ? d - LLR_CLAMP : v - LLR_CLAMP; d = LLR_CLAMP; }
                    else if (d < -LLR_CLAMP) { d = -LLR_CLAMP; }
                    r = exp(d); ri = exp(-d);
                    base += u;
                    dmax = fmax(dmax, fabs(d));          // a NaN likelihood leaves dmax alone; the NaN itself poisons the fit
                }
            }
            dst[l * 32] = r;
            dsti[l * 32] = ri;
        }
        rb[rb_off[warp] + ch * 32 + lane] = (packed && m < M) ? lm_r[m] + base : base;
    }
    double idm = 0.0;
    for (int64_t l = lane; l < L - 1; l += 32) idm = fmax(idm, invd[g0 + l]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
        idm = fmax(idm, __shfl_xor_sync(0xffffffffu, idm, o));
    }
    if (lane == 0) {
        rdesc[warp].llrmax = __double2float_ru(dmax);
        rdesc[warp].idmax = __double2float_ru(idm);
    }
}

// ---------------------------------------------------------------------------------------------------
// shared: stage the per-group potentials of one instance for groups [l0, l0+n)
// ---------------------------------------------------------------------------------------------------
struct TileRegs { double na[TILE / 32], nb[TILE / 32], id[TILE / 32], alpha[TILE / 32]; };
// first half: the loads of a tile (lanes strided over its groups) and the mask of groups with α_l ≥ 0 (bit j of m[j / 32])
__device__ __forceinline__ void tile_load(TileRegs& tr, int l0, int n, int lane, const double* __restrict__ na_g,
                                          const double* __restrict__ nb_g, const double* __restrict__ invd_g, double a, double b,
                                          unsigned& m0, unsigned& m1) {
    constexpr int K = TILE / 32;
    static_assert(K == 2, "two mask words");
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = lane + 32 * k, l = l0 + j;
        const bool in = j < n;
        tr.na[k] = in ? na_g[l] : 0.0;
        tr.nb[k] = in ? nb_g[l] : 0.0;
        tr.id[k] = (in && l > 0) ? invd_g[l - 1] : 0.0;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = lane + 32 * k;
        tr.alpha[k] = fma(b, tr.nb[k], a * tr.na[k]);
        const unsigned mk = __ballot_sync(0xffffffffu, j < n && __double2hiint(tr.alpha[k]) >= 0);
        if (k == 0) m0 = mk; else m1 = mk;
    }
}
// second half: the exponentials, and the potentials into shared memory
__device__ __forceinline__ void tile_store(SiteTile& st, const TileRegs& tr, int l0, int n, int lane, double cabs, bool bpos) {
    constexpr int K = TILE / 32;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = lane + 32 * k, l = l0 + j;
        if (j < n) {
            double e = cpn_exp_neg(-2.0 * fabs(tr.alpha[k]));
            bool pos = __double2hiint(tr.alpha[k]) >= 0;
            st.w[j] = make_double2(pos ? e : 1.0, pos ? 1.0 : e);
            // group 0 has no incoming coupling: t = 0 makes the generic step reproduce f = φ_0 from the all-ones message
            st.tn[j] = make_double2((l > 0) ? cpn_exp_neg(-2.0 * (cabs * tr.id[k])) : 0.0, tr.na[k]);
            st.bd[j] = make_double2(tr.nb[k], bpos ? tr.id[k] : -tr.id[k]);
        }
    }
}
__device__ __forceinline__ void fill_tile(SiteTile& st, int l0, int n, int lane, const double* __restrict__ na_g,
                                          const double* __restrict__ nb_g, const double* __restrict__ invd_g, double a, double b,
                                          double cabs, bool bpos) {
    TileRegs tr;
    unsigned m0, m1;
    tile_load(tr, l0, n, lane, na_g, nb_g, invd_g, a, b, m0, m1);
    tile_store(st, tr, l0, n, lane, cabs, bpos);
}

// ---------------------------------------------------------------------------------------------------
// E-step: ES = Σ_m ∇_ϕ log Zc_m(ϕ)   (≡ get_ess ∘ get_cond_exs_log)
// ---------------------------------------------------------------------------------------------------
// One forward sweep for the 32 reads of a chunk.  BPOS = (c ≥ 0).  (u,v) below are the source states that
// enter x'=+1 with weight 1 and t respectively; for c<0 the roles of same/opposite swap.
//
// The statistics are RATIOS of the tangent sums to the message sum, so the common scale of the eight state values is free.
// With potentials relative to their larger entry a step multiplies the larger state value by at most 2·max(1, r) and by at
// least t·min(1, r) (r = e^{LLR}, t = e^{−2|β|}): with B = max|LLR| + 2·max|β| the scale moves by at most e^{±(B + ln 2)} per
// group.  k_make_desc turns the region's bounds (RegionDesc::llrmax / idmax) and the instance's |c| into a rescale period:
// every second block of E_PF groups (B ≤ 40), every block (B ≤ 80) — 16 resp. 8 groups stay within e^{±650} — or, for anything
// wilder, the plain loop estep_chunk_slow with a rescale after every group.  Real and simulated nanopolish LLRs are a few
// tens at most, so the chain step itself carries no rescale arithmetic: 24 FP64 instructions per (read, group).
#ifndef E_PF_N
#define E_PF_N 8
#endif
constexpr int E_PF = E_PF_N;       // groups per block of emission rows (ring slot)
constexpr float E_BOUND_1 = 80.0f, E_BOUND_2 = 40.0f;     // bound on B for a rescale every block / every second block
constexpr int ITEM_MODE_SHIFT = 28;                       // ItemDesc::M bits 28-29: 0 slow path, 1 every block, 2 every second block
constexpr int ITEM_M_MASK = (1 << ITEM_MODE_SHIFT) - 1;

struct EState { double fp, fm, gap, gam, gbp, gbm, gcp, gcm; };

// One chain step for one read.  (u,v) are the source states entering x'=+1 with weight 1 and t; for c<0 the roles of
// aligned/opposite swap (BPOS=false).  SGN = 0: any α_l, both states multiplied by their relative potential (r rides on
// x'=+1).  SGN = +1 / −1: α_l ≥ 0 / α_l < 0 is known, the step is normalised by the potential of the heavier state, which
// therefore takes no multiplication at all; the other one is multiplied by κ = e^{−2|α_l|}·r^{∓1} (`r` is 1/r for SGN = +1).
template <bool BPOS, int SGN>
__device__ __forceinline__ void estep_site(EState& s, const SiteTile& st, int j, double r) {
    const double2 w = st.w[j], tn = st.tn[j], bd = st.bd[j];
    const double t = tn.x, na = tn.y, nb = bd.x, sinvd = bd.y;
    double u = BPOS ? s.fp : s.fm, v = BPOS ? s.fm : s.fp;
    double Gp = fma(t, v, u), Gm = fma(t, u, v);
    double ua = BPOS ? s.gap : s.gam, va = BPOS ? s.gam : s.gap;
    double ub = BPOS ? s.gbp : s.gbm, vb = BPOS ? s.gbm : s.gbp;
    double uc = BPOS ? s.gcp : s.gcm, vc = BPOS ? s.gcm : s.gcp;
    double Hap = fma(t, va, ua), Ham = fma(t, ua, va);
    double Hbp = fma(t, vb, ub), Hbm = fma(t, ub, vb);
    double Hcp = fma(t, vc, uc), Hcm = fma(t, uc, vc);
    double Dp = fma(2.0, u, -Gp), Dm = fma(2.0, v, -Gm);          // Σ_x (x x') ψ f, up to the sign carried by sinvd
    if (SGN == 0) {
        const double phiP = w.y * r, phiM = w.x;
        s.gap = phiP * fma(na, Gp, Hap);     s.gam = phiM * fma(-na, Gm, Ham);
        s.gbp = phiP * fma(nb, Gp, Hbp);     s.gbm = phiM * fma(-nb, Gm, Hbm);
        s.gcp = phiP * fma(sinvd, Dp, Hcp);  s.gcm = phiM * fma(sinvd, Dm, Hcm);
        s.fp = phiP * Gp;                    s.fm = phiM * Gm;
    } else if (SGN > 0) {
        const double k = w.x * r;
        s.gap = fma(na, Gp, Hap);            s.gam = k * fma(-na, Gm, Ham);
        s.gbp = fma(nb, Gp, Hbp);            s.gbm = k * fma(-nb, Gm, Hbm);
        s.gcp = fma(sinvd, Dp, Hcp);         s.gcm = k * fma(sinvd, Dm, Hcm);
        s.fp = Gp;                           s.fm = k * Gm;
    } else {
        const double k = w.y * r;
        s.gap = k * fma(na, Gp, Hap);        s.gam = fma(-na, Gm, Ham);
        s.gbp = k * fma(nb, Gp, Hbp);        s.gbm = fma(-nb, Gm, Hbm);
        s.gcp = k * fma(sinvd, Dp, Hcp);     s.gcm = fma(sinvd, Dm, Hcm);
        s.fp = k * Gp;                       s.fm = Gm;
    }
}
__device__ __forceinline__ void estep_rescale(EState& s) {
    const double sc = cpn_inv_pow2_of(s.fp + s.fm);
    s.fp *= sc; s.fm *= sc; s.gap *= sc; s.gam *= sc; s.gbp *= sc; s.gbm *= sc; s.gcp *= sc; s.gcm *= sc;
}

// Emission weights of a warp travel global → shared memory into a private ring: two slots of E_PF groups plus one slot for
// the chain tail.  Two transports:
//  * TMA (plain regions): the E_PF rows of a block are 2 KB contiguous in global memory, so lane 0 issues ONE bulk copy per
//    block (cp.async.bulk, completion on an mbarrier of the warp).  ncu showed the L1/shared-memory data pipe, not the FP64 pipe,
//    to be the busiest unit of this kernel (71 % vs 59 %): with per-lane cp.async every emission crossed it three times (L1
//    read, shared-memory write, shared-memory read); the bulk copy goes L2 → shared memory past it and leaves the one LDS.64.
//  * per-lane cp.async (views of the two-sample test, whose lanes read arbitrary pooled reads): 8 bytes per lane and group; a
//    lane only ever reads what it copied itself, so no warp synchronisation is needed.
// Either way the deep prefetch costs no registers (the register file is what limits the number of resident E-step warps).
struct EmisRing { double r[3][E_PF][32]; };

__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(mbar), "r"(parity) : "memory");
}
// lane 0 only: arm the barrier with the byte count and start the bulk copy of `bytes` (multiple of 16) contiguous bytes
__device__ __forceinline__ void tma_rows(unsigned dst, const double* __restrict__ src, unsigned bytes, unsigned mbar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the slot's earlier reads (generic proxy) precede the refill
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(mbar) : "memory");
}

// What the two transports need to know about the chunk of reads a warp is working on.
template <bool VIEW> struct EmisSrc;
template <> struct EmisSrc<false> {
    const double *p, *pi;      // row 0 of the chunk (lane 0) in the r and the 1/r plane
    unsigned mbar;             // shared-memory address of the warp's three barriers (8 bytes apart), one per ring slot
    unsigned ring;             // shared-memory address of the warp's ring
    unsigned phase;            // bit s: parity the next completion of slot s will have
};
template <> struct EmisSrc<true> {
    const double *p, *pi;      // this lane's read: its element of row 0 in either plane
};

// start the copy of rows [row0, row0 + nrows) of the plane chosen by `inv` into `slot` (nrows may be ≤ 0: nothing happens)
__device__ __forceinline__ void ering_issue(EmisRing& ring, int slot, EmisSrc<false>& es, bool inv, int row0, int nrows, int lane) {
    if (nrows > 0) {
        __syncwarp();                                    // every lane is done with what the slot held
        if (lane == 0)
            tma_rows(es.ring + (unsigned)slot * (unsigned)(E_PF * 256), (inv ? es.pi : es.p) + row0 * 32, (unsigned)nrows * 256u,
                     es.mbar + 8u * (unsigned)slot);
    }
}
__device__ __forceinline__ void ering_issue(EmisRing& ring, int slot, EmisSrc<true>& es, bool inv, int row0, int nrows, int lane) {
    const double* src = (inv ? es.pi : es.p) + (size_t)row0 * 32;
#pragma unroll
    for (int u = 0; u < E_PF; ++u) {
        if (u < nrows) {
            unsigned d = (unsigned)__cvta_generic_to_shared(&ring.r[slot][u][lane]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src + u * 32));
        }
    }
    asm volatile("cp.async.commit_group;");
}
// wait until the copy into `slot` has landed; `pending_after`: copies issued after it that may still be in flight (views)
template <int PENDING_AFTER>
__device__ __forceinline__ void ering_wait(EmisSrc<false>& es, int slot) {
    mbar_wait(es.mbar + 8u * (unsigned)slot, (es.phase >> slot) & 1u);
    es.phase ^= 1u << slot;
}
template <int PENDING_AFTER>
__device__ __forceinline__ void ering_wait(EmisSrc<true>&, int) {
    if (PENDING_AFTER == 0) asm volatile("cp.async.wait_group 0;");
    else asm volatile("cp.async.wait_group 1;");
}

// Forward sweep of the conditional chain for the 32 reads of a chunk, carrying the tangents (∂a,∂b,∂c).
// Groups are processed in blocks of E_PF: while one block of emission rows is consumed from the ring the next one is in
// flight.  A block whose groups all have α_l ≥ 0 (or all α_l < 0; 97 % of the blocks of the simulated genome, because ρ_l is
// small and a decides the sign) runs the matching specialised steps and reads 1/r (or r); a mixed block runs the general step.
// The state is rescaled by a power of two between blocks (rmask = 0: before every block, 1: before every second one).
template <bool BPOS, bool VIEW>
__device__ __forceinline__ void estep_chunk(SiteTile& st, EmisRing& ring, int L, int lane, const double* __restrict__ na_g,
                                            const double* __restrict__ nb_g, const double* __restrict__ invd_g, EmisSrc<VIEW>& es,
                                            double a, double b, double cabs, int rmask, double out[3]) {
    EState s{1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const int Lfull = L & ~(E_PF - 1);
    constexpr unsigned ALL = (1u << E_PF) - 1;
    // the chain tail (always the general step, so always r) is in flight from the start
    ering_issue(ring, 2, es, false, Lfull, L - Lfull, lane);
    int g = 0;                                          // index of the block of E_PF groups being consumed
    for (int l0 = 0; l0 < L; l0 += TILE) {
        const int n = min(TILE, L - l0);
        const int nfull = n & ~(E_PF - 1);
        TileRegs tr;
        unsigned m0, m1;
        tile_load(tr, l0, n, lane, na_g, nb_g, invd_g, a, b, m0, m1);
        // first block of the tile: its plane is known as soon as the signs are; the copy overlaps the exponentials below
        ering_issue(ring, g & 1, es, (m0 & ALL) == ALL, g * E_PF, nfull > 0 ? E_PF : 0, lane);
        __syncwarp();
        tile_store(st, tr, l0, n, lane, cabs, BPOS);
        __syncwarp();
#pragma unroll 1
        for (int j0 = 0; j0 < nfull; j0 += E_PF, ++g) {
            const unsigned bits = (((j0 & 32) ? m1 : m0) >> (j0 & 31)) & ALL;
            // next block of this tile in flight, then wait for block g
            const int j1 = j0 + E_PF;
            const unsigned bits1 = (((j1 & 32) ? m1 : m0) >> (j1 & 31)) & ALL;
            ering_issue(ring, (g + 1) & 1, es, bits1 == ALL, (g + 1) * E_PF, (j1 < nfull) ? E_PF : 0, lane);
            const int slot = g & 1;
            ering_wait<1>(es, slot);
            if (g > 0 && (g & rmask) == 0) estep_rescale(s);
            if (bits == ALL) {
#pragma unroll
                for (int u = 0; u < E_PF; ++u) estep_site<BPOS, 1>(s, st, j0 + u, ring.r[slot][u][lane]);
            } else if (bits == 0u) {
#pragma unroll
                for (int u = 0; u < E_PF; ++u) estep_site<BPOS, -1>(s, st, j0 + u, ring.r[slot][u][lane]);
            } else {
#pragma unroll
                for (int u = 0; u < E_PF; ++u) estep_site<BPOS, 0>(s, st, j0 + u, ring.r[slot][u][lane]);
            }
        }
        // tail of the chain (fewer than E_PF groups left; only the last tile has one)
        if (nfull < n) {
            ering_wait<0>(es, 2);
            if (g > 0) estep_rescale(s);
#pragma unroll 1
            for (int u = 0; u < n - nfull; ++u) estep_site<BPOS, 0>(s, st, nfull + u, ring.r[2][u][lane]);
        }
    }
    if (VIEW) asm volatile("cp.async.wait_group 0;");
    double inv = 1.0 / (s.fp + s.fm);
    out[0] = (s.gap + s.gam) * inv;
    out[1] = (s.gbp + s.gbm) * inv;
    out[2] = (s.gcp + s.gcm) * inv;
}

// The same sweep with a rescale after every group and plain loads: for instances whose |LLR| / |β| bound allows no unscaled
// run of groups (|LLR| is clamped at 300, so one group always fits the double range).
template <bool BPOS>
__device__ __noinline__ void estep_chunk_slow(SiteTile& st, int L, int lane, const double* __restrict__ na_g,
                                              const double* __restrict__ nb_g, const double* __restrict__ invd_g,
                                              const double* __restrict__ p, double a, double b, double cabs, double out[3]) {
    EState s{1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int l0 = 0; l0 < L; l0 += TILE) {
        const int n = min(TILE, L - l0);
        __syncwarp();
        fill_tile(st, l0, n, lane, na_g, nb_g, invd_g, a, b, cabs, BPOS);
        __syncwarp();
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            estep_site<BPOS, 0>(s, st, j, __ldg(p + (size_t)(l0 + j) * 32));
            estep_rescale(s);
        }
    }
    double inv = 1.0 / (s.fp + s.fm);
    out[0] = (s.gap + s.gam) * inv;
    out[1] = (s.gbp + s.gbm) * inv;
    out[2] = (s.gcp + s.gcm) * inv;
}

#ifndef E_BLOCKS
#define E_BLOCKS 5
#endif
#ifndef E_PERSIST
#define E_PERSIST 0                 // 1: resident grid, every warp walks the list with a block stride and prefetches its next instance
                                    // (measured 58-60 ms against 51 ms per 1e5 regions for one block per four instances: kept for experiments)
#endif
#ifndef E_PREFETCH
#define E_PREFETCH 1
#endif
#ifndef E_CARVEOUT
#define E_CARVEOUT 0
#endif
__device__ __forceinline__ void pf_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// One warp per (region, start) instance, lanes = reads.  The head of an instance is a chain of dependent loads (descriptor and ϕ →
// group tables → first rows of emissions) that the chain loop cannot start without; a warp therefore pulls the lines its NEXT
// instance will need into L1 while it still works on the current one — descriptor and ϕ at the start, the group tables and the
// first block of emission rows (both planes) when its chain ends, so that they travel during the reduction over reads.
template <bool VIEW>
__global__ void __launch_bounds__(E_WARPS * 32, E_BLOCKS) k_estep(int n_items, const ItemDesc* __restrict__ desc,
                                                                   const double* __restrict__ phil /*[3][cap] by list position*/, int64_t cap,
                                                                   RegionTables rt, int64_t NI, double* __restrict__ ES /*[3][NI]*/,
                                                                   const int* __restrict__ n_dev /*null, or the true item count*/) {
    __shared__ SiteTile tiles[E_WARPS];
    __shared__ EmisRing rings[E_WARPS];
    __shared__ longlong4 nextd[E_WARPS];          // {L, first group, first emission} of the warp's next instance
    __shared__ unsigned long long mbars[E_WARPS][4];   // per warp: one completion barrier per ring slot (TMA transport)
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    EmisSrc<VIEW> es;
    if constexpr (!VIEW) {
        es.mbar = (unsigned)__cvta_generic_to_shared(&mbars[w][0]);
        es.ring = (unsigned)__cvta_generic_to_shared(&rings[w].r[0][0][0]);
        es.phase = 0u;
        if (lane == 0) {
            mbar_init(es.mbar, 1u); mbar_init(es.mbar + 8u, 1u); mbar_init(es.mbar + 16u, 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    if (n_dev) n_items = __ldg(n_dev);
    const int stride = gridDim.x * E_WARPS;
    // Every block walks its own items (list positions ≡ 4·blockIdx + w mod stride) starting from a different round, so that the
    // blocks resident on an SM sit at different places of the length-sorted list: their instances then differ in length and the
    // latency-bound heads of some overlap the pipe-bound chain loops of the others instead of all 20 warps marching in step.
    const int rounds = E_PERSIST ? (n_items + stride - 1) / stride : 1;
    const int k0 = E_PERSIST ? (int)(((unsigned)blockIdx.x * 7919u) % (unsigned)max(rounds, 1)) : 0;
#pragma unroll 1
    for (int j = 0; j < rounds; ++j) {
        int k = k0 + j; if (k >= rounds) k -= rounds;
        const int item = k * stride + blockIdx.x * E_WARPS + w;
        if (item >= n_items) continue;
        int kn = k + 1; if (kn >= rounds) kn -= rounds;
        const int next = kn * stride + blockIdx.x * E_WARPS + w;
        const bool has_next = E_PERSIST && E_PREFETCH && j + 1 < rounds && next < n_items;
        // one round of independent loads: descriptor (two 16-byte vectors) and ϕ; and, for later, the next instance's descriptor
        const int4* dq = reinterpret_cast<const int4*>(desc + item);
        const int4 d0 = __ldg(dq);
        const longlong2 d1 = __ldg(reinterpret_cast<const longlong2*>(dq + 1));
        double a = phil[item], b = phil[cap + item], c = phil[2 * cap + item];
        if (has_next) {
            const int4* nq = reinterpret_cast<const int4*>(desc + next);
            const int4 n0 = __ldg(nq);
            const longlong2 n1 = __ldg(reinterpret_cast<const longlong2*>(nq + 1));
            if (lane < 3) pf_l1(phil + (size_t)lane * cap + next);
            if (lane == 0) nextd[w] = make_longlong4((long long)n0.y, n1.x, n1.y, 0);
        }
        const int inst = d0.x, L = d0.y, M = d0.z & ITEM_M_MASK, ridx = d0.w, mode = (d0.z >> ITEM_MODE_SHIFT) & 3;
        const int64_t g0 = d1.x;
        bool bpos = __double2hiint(c) >= 0;
        double cabs = fabs(c);
        const double* ew_r = rt.ew + d1.y;
        double acc0 = 0, acc1 = 0, acc2 = 0;
        int nchunk = (M + 31) >> 5;
        for (int ch = 0; ch < nchunk; ++ch) {
            double o[3];
            // this lane's read: lane of chunk ch, or — in a view — the pooled read the view's index row names
            int q = ch * 32 + lane;
            if (VIEW && ridx >= 0) q = __ldg(rt.ridx + ((size_t)(ridx + ch) << 5) + lane);
            const size_t po = (size_t)(q >> 5) * L * 32 + (q & 31);
            const double* p = ew_r + po;
            if (mode != 0) {
                // plain regions: row 0 of the chunk (the bulk copy moves whole rows); views: this lane's pooled read
                const size_t eo = VIEW ? po : (size_t)ch * L * 32;
                es.p = ew_r + eo;
                es.pi = rt.ewi + d1.y + eo;
                if (bpos) estep_chunk<true, VIEW>(tiles[w], rings[w], L, lane, rt.na + g0, rt.nb + g0, rt.invd + g0, es, a, b, cabs, mode - 1, o);
                else estep_chunk<false, VIEW>(tiles[w], rings[w], L, lane, rt.na + g0, rt.nb + g0, rt.invd + g0, es, a, b, cabs, mode - 1, o);
            } else {
                if (bpos) estep_chunk_slow<true>(tiles[w], L, lane, rt.na + g0, rt.nb + g0, rt.invd + g0, p, a, b, cabs, o);
                else estep_chunk_slow<false>(tiles[w], L, lane, rt.na + g0, rt.nb + g0, rt.invd + g0, p, a, b, cabs, o);
            }
            if (has_next && ch == nchunk - 1) {
                // the next instance's group tables (lanes 0-14: five 128-byte lines of each of na, nb, 1/d) and the first E_PF rows
                // of both emission planes (32 lines, one per lane) travel while this instance's reads are reduced
                __syncwarp();
                const longlong4 nd = nextd[w];
                const int Ln = (int)nd.x;
                if (lane < 15) {
                    const int tb = lane / 5, ln = lane - 5 * tb;
                    const double* base = (tb == 0 ? rt.na : (tb == 1 ? rt.nb : rt.invd)) + nd.y;
                    if (ln * 16 < Ln + 16) pf_l1(base + ln * 16);
                }
                const double* pl = ((lane & 16) ? rt.ewi : rt.ew) + nd.z + (size_t)((lane >> 1) & 7) * 32 + (lane & 1) * 16;
                if (((lane >> 1) & 7) < Ln) pf_l1(pl);
            }
            bool live = ch * 32 + lane < M;
            acc0 += cpn_warp_sum(live ? o[0] : 0.0);
            acc1 += cpn_warp_sum(live ? o[1] : 0.0);
            acc2 += cpn_warp_sum(live ? o[2] : 0.0);
        }
        if (lane == 0) {
            ES[inst] = acc0; ES[NI + inst] = acc1; ES[2 * NI + inst] = acc2;
        }
    }
}
// grid of an E-step launch over n items
static unsigned estep_grid(cpn_ctx* ctx, int64_t n) {
    const unsigned full = (unsigned)((n + E_WARPS - 1) / E_WARPS);
    if (!E_PERSIST) return full;
    if (ctx->estep_blocks <= 0) {          // what the device really keeps resident (registers, shared-memory carveout), not what was asked for
        int nb = 0;
        if (E_CARVEOUT) cudaFuncSetAttribute(k_estep<false>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_estep<false>, E_WARPS * 32, 0) != cudaSuccess || nb < 1) nb = 1;
        ctx->estep_blocks = nb;
    }
    return std::min<unsigned>(full, (unsigned)(ctx->sm_count * ctx->estep_blocks));
}

// Builds the position-indexed descriptors for a list of instances (items == null: identity list).
__global__ void k_make_desc(int n_items, const int* __restrict__ items, int n_init, const RegionDesc* __restrict__ rdesc,
                            const double* __restrict__ phi /*[3][NI]*/, int64_t NI, ItemDesc* __restrict__ desc, double* __restrict__ phil,
                            int64_t cap, const int* __restrict__ n_dev) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_dev) n_items = __ldg(n_dev);
    if (i >= n_items) return;
    int inst = items ? items[i] : i;
    RegionDesc rd = rdesc[inst / n_init];
    ItemDesc d;
    const double c = phi[2 * NI + inst];
    // rescale policy of the E-step for this instance (see estep_site); comparisons written so that NaN bounds or a NaN c choose slow
    const double B = (double)rd.llrmax + 2.0 * fabs(c) * (double)rd.idmax;
    const int mode = (B <= (double)E_BOUND_2) ? 2 : ((B <= (double)E_BOUND_1) ? 1 : 0);
    d.inst = inst; d.L = rd.L; d.M = rd.M | (mode << ITEM_MODE_SHIFT); d.ridx = rd.ridx; d.g0 = rd.g0; d.ew_off = rd.ew_off;
    desc[i] = d;
    phil[i] = phi[inst]; phil[cap + i] = phi[NI + inst]; phil[2 * cap + i] = c;
}

// ---------------------------------------------------------------------------------------------------
// E-step, two EM instances per warp (the default for the analytic M-step; k_estep above stays as the one-instance form)
// ---------------------------------------------------------------------------------------------------
// k_estep's time follows its executed issue slots, and a third of them are not chain arithmetic: the head and tail of an
// instance (descriptor, group potentials, reductions), the per-block emission transport and three warp-uniform LDS.128 per
// chain step.  All of that is per WARP, so the pair form halves it per instance: lanes 0-15 work for one start of a region and
// lanes 16-31 for another start of the SAME region, every lane carrying two reads (slots i and i+16 of the 32-read chunk).  The
// two instances share the region's emission rows (one bulk copy per block for both), the block bookkeeping and the loop; the
// potentials of a step are one LDS.128 triple per lane for two chain steps.  Both halves run one instruction stream, so
//  * only starts whose coupling c has the same sign are paired (the BPOS specialisation is per warp);
//  * a block of E_PF groups runs a sign-specialised body only when BOTH instances have that sign of α_l on all its groups,
//    otherwise the general body (25 instead of 21 FP64 instructions per step, r plane).
// Pairs are rebuilt every EM iteration from the active list (k_make_pairs): within a region's run of active starts, starts of
// one sign class are paired in list order; an odd one out runs with its second half idle.  Pairing only looks at the region's
// own starts, so results do not depend on how a batch is chunked or sharded.  The sum over reads is the same tree as
// cpn_warp_sum (first i with i+16, then the xor butterfly 8,4,2,1).
struct __align__(16) PairDesc { int instA, instB, L, M; int64_t g0, ew_off; int ridx, pad0, pad1, pad2; };
static_assert(sizeof(PairDesc) == 48, "PairDesc layout");
#ifndef EP_BLOCKS
#define EP_BLOCKS 4
#endif
#ifndef EP_UNROLL
#define EP_UNROLL 2                 // groups per trip of a block body (see estep_pair_chunk)
#endif
constexpr int EP_UNR = EP_UNROLL;

// potentials of the two instances of a pair; rows padded by one entry so that the two halves' broadcast reads of the same group
// fall into different banks
struct PairTile { double2 w[2][TILE + 1], tn[2][TILE + 1], bd[2][TILE + 1]; };
#ifdef EP_STATS          // diagnostic build: how many blocks run each body, how many pairs are singletons
__device__ unsigned long long ep_stats[8];
#define EP_COUNT(k) do { if (lane == 0) atomicAdd(&ep_stats[k], 1ull); } while (0)
#else
#define EP_COUNT(k) do { } while (0)
#endif
constexpr size_t EP_SMEM = (size_t)E_WARPS * (sizeof(PairTile) + sizeof(EmisRing)) + (size_t)E_WARPS * 4 * sizeof(unsigned long long);

// estep_site on explicit operands (same arithmetic, same order)
template <bool BPOS, int SGN>
__device__ __forceinline__ void estep_core(EState& s, const double2 w, const double2 tn, const double2 bd, double r) {
    const double t = tn.x, na = tn.y, nb = bd.x, sinvd = bd.y;
    double u = BPOS ? s.fp : s.fm, v = BPOS ? s.fm : s.fp;
    double Gp = fma(t, v, u), Gm = fma(t, u, v);
    double ua = BPOS ? s.gap : s.gam, va = BPOS ? s.gam : s.gap;
    double ub = BPOS ? s.gbp : s.gbm, vb = BPOS ? s.gbm : s.gbp;
    double uc = BPOS ? s.gcp : s.gcm, vc = BPOS ? s.gcm : s.gcp;
    double Hap = fma(t, va, ua), Ham = fma(t, ua, va);
    double Hbp = fma(t, vb, ub), Hbm = fma(t, ub, vb);
    double Hcp = fma(t, vc, uc), Hcm = fma(t, uc, vc);
    double Dp = fma(2.0, u, -Gp), Dm = fma(2.0, v, -Gm);
    if (SGN == 0) {
        const double phiP = w.y * r, phiM = w.x;
        s.gap = phiP * fma(na, Gp, Hap);     s.gam = phiM * fma(-na, Gm, Ham);
        s.gbp = phiP * fma(nb, Gp, Hbp);     s.gbm = phiM * fma(-nb, Gm, Hbm);
        s.gcp = phiP * fma(sinvd, Dp, Hcp);  s.gcm = phiM * fma(sinvd, Dm, Hcm);
        s.fp = phiP * Gp;                    s.fm = phiM * Gm;
    } else if (SGN > 0) {
        const double k = w.x * r;
        s.gap = fma(na, Gp, Hap);            s.gam = k * fma(-na, Gm, Ham);
        s.gbp = fma(nb, Gp, Hbp);            s.gbm = k * fma(-nb, Gm, Hbm);
        s.gcp = fma(sinvd, Dp, Hcp);         s.gcm = k * fma(sinvd, Dm, Hcm);
        s.fp = Gp;                           s.fm = k * Gm;
    } else {
        const double k = w.y * r;
        s.gap = k * fma(na, Gp, Hap);        s.gam = fma(-na, Gm, Ham);
        s.gbp = k * fma(nb, Gp, Hbp);        s.gbm = fma(-nb, Gm, Hbm);
        s.gcp = k * fma(sinvd, Dp, Hcp);     s.gcm = fma(sinvd, Dm, Hcm);
        s.fp = k * Gp;                       s.fm = Gm;
    }
}

// Potentials of groups [l0, l0+n) for both instances: half h of the warp (16 lanes strided over the groups) fills its own rows.
// mP / mN: bit j set when BOTH instances have α ≥ 0 / α < 0 at group l0 + j.
struct PairTileRegs { double na[TILE / 16], nb[TILE / 16], id[TILE / 16], alpha[TILE / 16]; };
__device__ __forceinline__ void ptile_load(PairTileRegs& tr, int l0, int n, int i, const double* __restrict__ na_g,
                                           const double* __restrict__ nb_g, const double* __restrict__ invd_g, double a, double b,
                                           unsigned long long& mP, unsigned long long& mN) {
    constexpr int K = TILE / 16;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = i + 16 * k, l = l0 + j;
        const bool in = j < n;
        tr.na[k] = in ? na_g[l] : 0.0;
        tr.nb[k] = in ? nb_g[l] : 0.0;
        tr.id[k] = (in && l > 0) ? invd_g[l - 1] : 0.0;
    }
    mP = 0ull; mN = 0ull;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = i + 16 * k;
        tr.alpha[k] = fma(b, tr.nb[k], a * tr.na[k]);
        const unsigned bal = __ballot_sync(0xffffffffu, j < n && __double2hiint(tr.alpha[k]) >= 0);
        mP |= (unsigned long long)(bal & (bal >> 16) & 0xffffu) << (16 * k);
        mN |= (unsigned long long)(~(bal | (bal >> 16)) & 0xffffu) << (16 * k);
    }
}
// The tile's exponentials without the range tests of cpn_exp_neg (cpn_exp_neg_clamped: 7 instructions and a branch less per
// exponential): a start with a non-finite ϕ, whose NaN the tests used to carry into ES, is caught by pair_poison instead.
#ifndef EP_EXP_CLAMPED
#define EP_EXP_CLAMPED 1
#endif
#if EP_EXP_CLAMPED
#define EP_EXP cpn_exp_neg_clamped
#else
#define EP_EXP cpn_exp_neg
#endif
// ES of a start whose ϕ is not finite is NaN (the M-step then reports the numerical failure, em_algorithm.jl:160-172)
__device__ __forceinline__ void pair_poison(double a, double b, double c, double& acc0, double& acc1, double& acc2) {
#if EP_EXP_CLAMPED
    if (!(isfinite(a) && isfinite(b) && isfinite(c))) { acc0 = cpn_nan(); acc1 = cpn_nan(); acc2 = cpn_nan(); }
#endif
}
__device__ __forceinline__ void ptile_store(PairTile& pt, const PairTileRegs& tr, int l0, int n, int h, int i, double cabs, bool bpos) {
    constexpr int K = TILE / 16;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = i + 16 * k, l = l0 + j;
        if (j < n) {
            const double e = EP_EXP(-2.0 * fabs(tr.alpha[k]));
            const bool pos = __double2hiint(tr.alpha[k]) >= 0;
            pt.w[h][j] = make_double2(pos ? e : 1.0, pos ? 1.0 : e);
            pt.tn[h][j] = make_double2((l > 0) ? EP_EXP(-2.0 * (cabs * tr.id[k])) : 0.0, tr.na[k]);
            pt.bd[h][j] = make_double2(tr.nb[k], bpos ? tr.id[k] : -tr.id[k]);
        }
    }
}

// ring transport for a pair: with the per-lane copies of a view a lane reads what ANOTHER lane copied (slot i+16, or i from the
// upper half), so the warp synchronises before a slot is refilled and after a copy has landed; the bulk-copy path already does
template <bool VIEW>
__device__ __forceinline__ void pring_issue(EmisRing& ring, int slot, EmisSrc<VIEW>& es, bool inv, int row0, int nrows, int lane) {
    if (VIEW) __syncwarp();
    ering_issue(ring, slot, es, inv, row0, nrows, lane);
}
template <int PENDING_AFTER, bool VIEW>
__device__ __forceinline__ void pring_wait(EmisSrc<VIEW>& es, int slot) {
    ering_wait<PENDING_AFTER>(es, slot);
    if (VIEW) __syncwarp();
}

// EP_UNR groups × 2 reads per trip of a block body.  The bodies are NOT unrolled over the whole block: six of them (three sign
// cases × the sign of c) are live on an SM at once, and fully unrolled (8 groups × 2 reads ≈ 6 KB each, next to two copies of the
// head) they fell out of the 32 KB instruction cache as soon as both signs of c were present — ncu: 75-85 % of the warp-stall
// samples of the bodies were "no instruction", 6.3 instead of 2.9 ms per launch.  For the same reason everything but the bodies
// exists once, with the sign of c as a run-time value.
template <bool BPOS, int SGN>
__device__ __forceinline__ void pair_block(EState& s0, EState& s1, const double2* __restrict__ W, const double2* __restrict__ TN,
                                           const double2* __restrict__ BD, const double* __restrict__ r /*this lane's first read, row 0*/,
                                           int rows) {
#pragma unroll EP_UNR
    for (int u = 0; u < rows; ++u) {
        const double2 w = W[u], tn = TN[u], bd = BD[u];
        estep_core<BPOS, SGN>(s0, w, tn, bd, r[u * 32]);
        estep_core<BPOS, SGN>(s1, w, tn, bd, r[u * 32 + 16]);
    }
}

template <bool VIEW>
__device__ __forceinline__ void estep_pair_chunk(PairTile& pt, EmisRing& ring, int L, int lane, const double* __restrict__ na_g,
                                                 const double* __restrict__ nb_g, const double* __restrict__ invd_g, EmisSrc<VIEW>& es,
                                                 double a, double b, double cabs, bool bpos, int rmask, double out0[3], double out1[3]) {
    const int h = lane >> 4, i = lane & 15;
    const double2* __restrict__ W = pt.w[h];
    const double2* __restrict__ TN = pt.tn[h];
    const double2* __restrict__ BD = pt.bd[h];
    EState s0{1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    EState s1 = s0;
    const int Lfull = L & ~(E_PF - 1);
    constexpr unsigned ALL = (1u << E_PF) - 1;
    pring_issue<VIEW>(ring, 2, es, false, Lfull, L - Lfull, lane);
    int g = 0;
    for (int l0 = 0; l0 < L; l0 += TILE) {
        const int n = min(TILE, L - l0);
        const int nfull = n & ~(E_PF - 1);
        PairTileRegs tr;
        unsigned long long mP, mN;
        ptile_load(tr, l0, n, i, na_g, nb_g, invd_g, a, b, mP, mN);
        pring_issue<VIEW>(ring, g & 1, es, ((unsigned)mP & ALL) == ALL, g * E_PF, nfull > 0 ? E_PF : 0, lane);
        __syncwarp();
        ptile_store(pt, tr, l0, n, h, i, cabs, bpos);
        __syncwarp();
#pragma unroll 1
        for (int j0 = 0; j0 < nfull; j0 += E_PF, ++g) {
            const unsigned bp = (unsigned)(mP >> j0) & ALL, bn = (unsigned)(mN >> j0) & ALL;
            const int j1 = j0 + E_PF;
            const unsigned bp1 = (unsigned)(mP >> (j1 & (TILE - 1))) & ALL;
            pring_issue<VIEW>(ring, (g + 1) & 1, es, bp1 == ALL, (g + 1) * E_PF, (j1 < nfull) ? E_PF : 0, lane);
            const int slot = g & 1;
            pring_wait<1, VIEW>(es, slot);
            if (g > 0 && (g & rmask) == 0) { estep_rescale(s0); estep_rescale(s1); }
            EP_COUNT(bp == ALL ? 0 : (bn == ALL ? 1 : 2));
            const double* r = &ring.r[slot][0][i];
            if (bpos) {
                if (bp == ALL) pair_block<true, 1>(s0, s1, W + j0, TN + j0, BD + j0, r, E_PF);
                else if (bn == ALL) pair_block<true, -1>(s0, s1, W + j0, TN + j0, BD + j0, r, E_PF);
                else pair_block<true, 0>(s0, s1, W + j0, TN + j0, BD + j0, r, E_PF);
            } else {
                if (bp == ALL) pair_block<false, 1>(s0, s1, W + j0, TN + j0, BD + j0, r, E_PF);
                else if (bn == ALL) pair_block<false, -1>(s0, s1, W + j0, TN + j0, BD + j0, r, E_PF);
                else pair_block<false, 0>(s0, s1, W + j0, TN + j0, BD + j0, r, E_PF);
            }
        }
        if (nfull < n) {
            pring_wait<0, VIEW>(es, 2);
            if (g > 0) { estep_rescale(s0); estep_rescale(s1); }
            const double* r = &ring.r[2][0][i];
#pragma unroll 1
            for (int u = 0; u < n - nfull; ++u) {
                const double2 w = W[nfull + u], tn = TN[nfull + u], bd = BD[nfull + u];
                if (bpos) { estep_core<true, 0>(s0, w, tn, bd, r[u * 32]); estep_core<true, 0>(s1, w, tn, bd, r[u * 32 + 16]); }
                else { estep_core<false, 0>(s0, w, tn, bd, r[u * 32]); estep_core<false, 0>(s1, w, tn, bd, r[u * 32 + 16]); }
            }
        }
    }
    if (VIEW) asm volatile("cp.async.wait_group 0;");
    const double inv0 = 1.0 / (s0.fp + s0.fm), inv1 = 1.0 / (s1.fp + s1.fm);
    out0[0] = (s0.gap + s0.gam) * inv0; out0[1] = (s0.gbp + s0.gbm) * inv0; out0[2] = (s0.gcp + s0.gcm) * inv0;
    out1[0] = (s1.gap + s1.gam) * inv1; out1[1] = (s1.gbp + s1.gbm) * inv1; out1[2] = (s1.gcp + s1.gcm) * inv1;
}

// rescale after every group, plain loads (pairs whose |LLR| / |β| bound allows no unscaled run of groups)
template <bool BPOS>
__device__ __noinline__ void estep_pair_slow(PairTile& pt, int L, int lane, const double* __restrict__ na_g, const double* __restrict__ nb_g,
                                             const double* __restrict__ invd_g, const double* __restrict__ p0, const double* __restrict__ p1,
                                             double a, double b, double cabs, double out0[3], double out1[3]) {
    const int h = lane >> 4, i = lane & 15;
    EState s0{1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    EState s1 = s0;
    for (int l0 = 0; l0 < L; l0 += TILE) {
        const int n = min(TILE, L - l0);
        PairTileRegs tr;
        unsigned long long mP, mN;
        __syncwarp();
        ptile_load(tr, l0, n, i, na_g, nb_g, invd_g, a, b, mP, mN);
        ptile_store(pt, tr, l0, n, h, i, cabs, BPOS);
        __syncwarp();
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            const double2 w = pt.w[h][j], tn = pt.tn[h][j], bd = pt.bd[h][j];
            estep_core<BPOS, 0>(s0, w, tn, bd, __ldg(p0 + (size_t)(l0 + j) * 32));
            estep_rescale(s0);
            estep_core<BPOS, 0>(s1, w, tn, bd, __ldg(p1 + (size_t)(l0 + j) * 32));
            estep_rescale(s1);
        }
    }
    const double inv0 = 1.0 / (s0.fp + s0.fm), inv1 = 1.0 / (s1.fp + s1.fm);
    out0[0] = (s0.gap + s0.gam) * inv0; out0[1] = (s0.gbp + s0.gbm) * inv0; out0[2] = (s0.gcp + s0.gcm) * inv0;
    out1[0] = (s1.gap + s1.gam) * inv1; out1[1] = (s1.gbp + s1.gbm) * inv1; out1[2] = (s1.gcp + s1.gcm) * inv1;
}

template <bool VIEW>
__global__ void __launch_bounds__(E_WARPS * 32, EP_BLOCKS) k_estep_pair(const int* __restrict__ n_pairs_dev, const PairDesc* __restrict__ pdesc,
                                                                         const double* __restrict__ phil2 /*[6][cap] by pair position*/,
                                                                         int64_t cap, RegionTables rt, int64_t NI, double* __restrict__ ES) {
    extern __shared__ __align__(128) unsigned char ep_smem[];
    PairTile* tiles = reinterpret_cast<PairTile*>(ep_smem);
    EmisRing* rings = reinterpret_cast<EmisRing*>(ep_smem + (size_t)E_WARPS * sizeof(PairTile));
    unsigned long long* mbars = reinterpret_cast<unsigned long long*>(ep_smem + (size_t)E_WARPS * (sizeof(PairTile) + sizeof(EmisRing)));
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, h = lane >> 4, i = lane & 15;
    EmisSrc<VIEW> es;
    if constexpr (!VIEW) {
        es.mbar = (unsigned)__cvta_generic_to_shared(&mbars[4 * w]);
        es.ring = (unsigned)__cvta_generic_to_shared(&rings[w].r[0][0][0]);
        es.phase = 0u;
        if (lane == 0) {
            mbar_init(es.mbar, 1u); mbar_init(es.mbar + 8u, 1u); mbar_init(es.mbar + 16u, 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    // pairs with c ≥ 0 fill the list from the front, pairs with c < 0 from the back (k_make_pairs): the launch runs the former
    // first and the latter together at its end, so that an SM mostly executes the bodies of ONE sign of c at a time
    const int n_pos = __ldg(n_pairs_dev), n_pairs = n_pos + __ldg(n_pairs_dev + 1);
#pragma unroll 1
    for (int pq = blockIdx.x * E_WARPS + w; pq < n_pairs; pq += gridDim.x * E_WARPS) {
        const int p = pq < n_pos ? pq : (int)cap - 1 - (pq - n_pos);
        const int4* dq = reinterpret_cast<const int4*>(pdesc + p);
        const int4 d0 = __ldg(dq);
        const longlong2 d1 = __ldg(reinterpret_cast<const longlong2*>(dq + 1));
        const int ridx = __ldg(reinterpret_cast<const int*>(dq + 2));
        const double a = phil2[(size_t)(3 * h) * cap + p], b = phil2[(size_t)(3 * h + 1) * cap + p], c = phil2[(size_t)(3 * h + 2) * cap + p];
        const int L = d0.z, M = d0.w & ITEM_M_MASK, mode = (d0.w >> ITEM_MODE_SHIFT) & 3;
        const int64_t g0 = d1.x;
        // both starts of a pair have the same sign of c (k_make_pairs); lane 0's copy decides for the warp
        const bool bpos = __shfl_sync(0xffffffffu, (int)(__double2hiint(c) >= 0), 0) != 0;
        const double cabs = fabs(c);
        const double* ew_r = rt.ew + d1.y;
        double acc0 = 0, acc1 = 0, acc2 = 0;
        const int nchunk = (M + 31) >> 5;
        for (int ch = 0; ch < nchunk; ++ch) {
            double o0[3], o1[3];
            int q = ch * 32 + lane;
            if (VIEW && ridx >= 0) q = __ldg(rt.ridx + ((size_t)(ridx + ch) << 5) + lane);
            const size_t po = (size_t)(q >> 5) * L * 32 + (q & 31);
            if (mode != 0) {
                const size_t eo = VIEW ? po : (size_t)ch * L * 32;
                es.p = ew_r + eo;
                es.pi = rt.ewi + d1.y + eo;
                estep_pair_chunk<VIEW>(tiles[w], rings[w], L, lane, rt.na + g0, rt.nb + g0, rt.invd + g0, es, a, b, cabs, bpos, mode - 1, o0, o1);
            } else {
                const int q0 = __shfl_sync(0xffffffffu, q, i), q1 = __shfl_sync(0xffffffffu, q, i + 16);
                const double* p0 = ew_r + (size_t)(q0 >> 5) * L * 32 + (q0 & 31);
                const double* p1 = ew_r + (size_t)(q1 >> 5) * L * 32 + (q1 & 31);
                if (bpos) estep_pair_slow<true>(tiles[w], L, lane, rt.na + g0, rt.nb + g0, rt.invd + g0, p0, p1, a, b, cabs, o0, o1);
                else estep_pair_slow<false>(tiles[w], L, lane, rt.na + g0, rt.nb + g0, rt.invd + g0, p0, p1, a, b, cabs, o0, o1);
            }
            const bool live0 = ch * 32 + i < M, live1 = ch * 32 + 16 + i < M;
            double v0 = (live0 ? o0[0] : 0.0) + (live1 ? o1[0] : 0.0);
            double v1 = (live0 ? o0[1] : 0.0) + (live1 ? o1[1] : 0.0);
            double v2 = (live0 ? o0[2] : 0.0) + (live1 ? o1[2] : 0.0);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                v0 += __shfl_xor_sync(0xffffffffu, v0, o);
                v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                v2 += __shfl_xor_sync(0xffffffffu, v2, o);
            }
            acc0 += v0; acc1 += v1; acc2 += v2;
        }
        EP_COUNT(d0.y < 0 ? 3 : 4);
        if (mode == 0) EP_COUNT(5);
        const int inst = h ? d0.y : d0.x;
        pair_poison(a, b, c, acc0, acc1, acc2);
        if (i == 0 && inst >= 0) { ES[inst] = acc0; ES[NI + inst] = acc1; ES[2 * NI + inst] = acc2; }
    }
}

// ---------------------------------------------------------------------------------------------------
// The paired E-step as a resident grid with a work queue and a head that is prefetched (plain regions)
// ---------------------------------------------------------------------------------------------------
// ncu (source page) on k_estep_pair: about a third of a warp's life is the HEAD of a pair — three dependent round trips to L2
// (descriptor and ϕ → group tables → first emission rows) during which it issues nothing — and with 16 warps per SM the FP64 pipe
// idles 37 % of the time.  Here every warp takes pairs from a queue (one atomic per pair: whoever is free takes the next one, so
// the warps of an SM may progress at different rates without idling, which the static walk of E_PERSIST could not offer) and
// brings its NEXT pair's descriptor, ϕ and group tables into shared memory with cp.async while it works on the current one:
//   loop top        descriptor and ϕ of the current pair: shared memory → registers; take the next pair from the queue, start the
//                   copy of its descriptor and ϕ (group A)
//   tile 0          potentials from the staged tables (shared memory, not global); first emission rows by TMA as before
//   after tile 0    wait for A, start the copy of the next pair's group tables (group B) — it lands during the chain loop
//   chain, epilogue as k_estep_pair
// What is left exposed per pair is one round trip (the first emission rows), overlapped with the exponentials of the tile.
struct __align__(16) PairStage { int4 d[3]; double phi[6]; double pad[2]; double na[TILE], nb[TILE], id[TILE]; };
constexpr size_t EPW_SMEM = (size_t)E_WARPS * (sizeof(PairTile) + sizeof(EmisRing) + sizeof(PairStage)) + (size_t)E_WARPS * 4 * sizeof(unsigned long long);

__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
// descriptor (three 16-byte parts, lanes 0-2) and ϕ of both instances (lanes 3-8) of the pair at list position p
__device__ __forceinline__ void stage_desc(PairStage& st, const PairDesc* __restrict__ pdesc, const double* __restrict__ phil2, int64_t cap,
                                           int p, int lane) {
    if (lane < 3) cp_async16(&st.d[lane], reinterpret_cast<const int4*>(pdesc + p) + lane);
    else if (lane < 9) cp_async8(&st.phi[lane - 3], phil2 + (size_t)(lane - 3) * cap + p);
    asm volatile("cp.async.commit_group;" ::: "memory");
}
// group tables of the first tile of a region: N_l, N_l ρ_l, and 1/d of the coupling INTO group l (0 for l = 0)
__device__ __forceinline__ void stage_tables(PairStage& st, const RegionTables& rt, int64_t g0, int L, int lane) {
    const int n = min(L, TILE);
    for (int j = lane; j < n; j += 32) {
        cp_async8(&st.na[j], rt.na + g0 + j);
        cp_async8(&st.nb[j], rt.nb + g0 + j);
        if (j > 0) cp_async8(&st.id[j], rt.invd + g0 + j - 1);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void ptile_load_staged(PairTileRegs& tr, const PairStage& st, int n, int i, double a, double b,
                                                  unsigned long long& mP, unsigned long long& mN) {
    constexpr int K = TILE / 16;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = i + 16 * k;
        const bool in = j < n;
        tr.na[k] = in ? st.na[j] : 0.0;
        tr.nb[k] = in ? st.nb[j] : 0.0;
        tr.id[k] = (in && j > 0) ? st.id[j] : 0.0;
    }
    mP = 0ull; mN = 0ull;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = i + 16 * k;
        tr.alpha[k] = fma(b, tr.nb[k], a * tr.na[k]);
        const unsigned bal = __ballot_sync(0xffffffffu, j < n && __double2hiint(tr.alpha[k]) >= 0);
        mP |= (unsigned long long)(bal & (bal >> 16) & 0xffffu) << (16 * k);
        mN |= (unsigned long long)(~(bal | (bal >> 16)) & 0xffffu) << (16 * k);
    }
}

// estep_pair_chunk for plain regions with the first tile's tables taken from the stage when `staged`; `next` ≥ 0: once the first
// tile has been consumed, the next pair's descriptor (group A, in flight since the loop top) is awaited and its tables are requested
#ifndef EP_RUNPTR
#define EP_RUNPTR 1
#endif
// one full block of E_PF rows (E_PF · 256 contiguous bytes at src) into the ring slot at shared-memory address dst
__device__ __forceinline__ void ering_issue_at(unsigned dst, unsigned mbar, const char* src, int lane) {
    __syncwarp();                                        // every lane is done with what the slot held
    if (lane == 0) tma_rows(dst, reinterpret_cast<const double*>(src), (unsigned)(E_PF * 256), mbar);
}
__device__ __forceinline__ void estep_pair_chunk_ws(PairTile& pt, EmisRing& ring, PairStage& stg, const RegionTables& rt, int L, int lane,
                                                    const double* __restrict__ na_g, const double* __restrict__ nb_g,
                                                    const double* __restrict__ invd_g, EmisSrc<false>& es, double a, double b, double cabs,
                                                    bool bpos, int rmask, bool staged, bool fetch_next, double out0[3], double out1[3]) {
    const int h = lane >> 4, i = lane & 15;
    const double2* __restrict__ W = pt.w[h];
    const double2* __restrict__ TN = pt.tn[h];
    const double2* __restrict__ BD = pt.bd[h];
    EState s0{1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    EState s1 = s0;
    const int Lfull = L & ~(E_PF - 1);
    constexpr unsigned ALL = (1u << E_PF) - 1;
    ering_issue(ring, 2, es, false, Lfull, L - Lfull, lane);
    // the full blocks' rows: a running pointer into the r plane (the 1/r plane lies dpi bytes further) instead of addresses rebuilt
    // from the chunk's offsets by the one issuing lane
#if EP_RUNPTR
    const char* nxt = reinterpret_cast<const char*>(es.p);
    const long long dpi = reinterpret_cast<const char*>(rt.ewi) - reinterpret_cast<const char*>(rt.ew);
#endif
    int g = 0;
    for (int l0 = 0; l0 < L; l0 += TILE) {
        const int n = min(TILE, L - l0);
        const int nfull = n & ~(E_PF - 1);
        PairTileRegs tr;
        unsigned long long mP, mN;
        if (staged && l0 == 0) ptile_load_staged(tr, stg, n, i, a, b, mP, mN);
        else ptile_load(tr, l0, n, i, na_g, nb_g, invd_g, a, b, mP, mN);
#if EP_RUNPTR
        if (nfull > 0) {
            ering_issue_at(es.ring + (unsigned)(g & 1) * (unsigned)(E_PF * 256), es.mbar + 8u * (unsigned)(g & 1),
                           nxt + ((((unsigned)mP & ALL) == ALL) ? dpi : 0ll), lane);
            nxt += E_PF * 256;
        }
#else
        ering_issue(ring, g & 1, es, ((unsigned)mP & ALL) == ALL, g * E_PF, nfull > 0 ? E_PF : 0, lane);
#endif
        __syncwarp();
        ptile_store(pt, tr, l0, n, h, i, cabs, bpos);
        __syncwarp();
        if (fetch_next && l0 == 0) {
            // the staged tables are in registers and in the tile now: the stage may take the next pair's
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            const int4 nd0 = stg.d[0];
            const longlong2 nd1 = *reinterpret_cast<const longlong2*>(&stg.d[1]);
            stage_tables(stg, rt, nd1.x, nd0.z, lane);
        }
#pragma unroll 1
        for (int j0 = 0; j0 < nfull; j0 += E_PF, ++g) {
            const unsigned bp = (unsigned)(mP >> j0) & ALL, bn = (unsigned)(mN >> j0) & ALL;
            const int j1 = j0 + E_PF;
            const unsigned bp1 = (unsigned)(mP >> (j1 & (TILE - 1))) & ALL;
#if EP_RUNPTR
            if (j1 < nfull) {
                ering_issue_at(es.ring + (unsigned)((g + 1) & 1) * (unsigned)(E_PF * 256), es.mbar + 8u * (unsigned)((g + 1) & 1),
                               nxt + (bp1 == ALL ? dpi : 0ll), lane);
                nxt += E_PF * 256;
            }
#else
            ering_issue(ring, (g + 1) & 1, es, bp1 == ALL, (g + 1) * E_PF, (j1 < nfull) ? E_PF : 0, lane);
#endif
            const int slot = g & 1;
            ering_wait<1>(es, slot);
            if (g > 0 && (g & rmask) == 0) { estep_rescale(s0); estep_rescale(s1); }
            const double* r = &ring.r[slot][0][i];
            if (bpos) {
                if (bp == ALL) pair_block<true, 1>(s0, s1, W + j0, TN + j0, BD + j0, r, E_PF);
                else if (bn == ALL) pair_block<true, -1>(s0, s1, W + j0, TN + j0, BD + j0, r, E_PF);
                else pair_block<true, 0>(s0, s1, W + j0, TN + j0, BD + j0, r, E_PF);
            } else {
                if (bp == ALL) pair_block<false, 1>(s0, s1, W + j0, TN + j0, BD + j0, r, E_PF);
                else if (bn == ALL) pair_block<false, -1>(s0, s1, W + j0, TN + j0, BD + j0, r, E_PF);
                else pair_block<false, 0>(s0, s1, W + j0, TN + j0, BD + j0, r, E_PF);
            }
        }
        if (nfull < n) {
            ering_wait<0>(es, 2);
            if (g > 0) { estep_rescale(s0); estep_rescale(s1); }
            const double* r = &ring.r[2][0][i];
#pragma unroll 1
            for (int u = 0; u < n - nfull; ++u) {
                const double2 w = W[nfull + u], tn = TN[nfull + u], bd = BD[nfull + u];
                if (bpos) { estep_core<true, 0>(s0, w, tn, bd, r[u * 32]); estep_core<true, 0>(s1, w, tn, bd, r[u * 32 + 16]); }
                else { estep_core<false, 0>(s0, w, tn, bd, r[u * 32]); estep_core<false, 0>(s1, w, tn, bd, r[u * 32 + 16]); }
            }
        }
    }
    const double inv0 = 1.0 / (s0.fp + s0.fm), inv1 = 1.0 / (s1.fp + s1.fm);
    out0[0] = (s0.gap + s0.gam) * inv0; out0[1] = (s0.gbp + s0.gbm) * inv0; out0[2] = (s0.gcp + s0.gcm) * inv0;
    out1[0] = (s1.gap + s1.gam) * inv1; out1[1] = (s1.gbp + s1.gbm) * inv1; out1[2] = (s1.gcp + s1.gcm) * inv1;
}

// cnt: {pairs with c ≥ 0, pairs with c < 0, queue head (zero at launch), unused}
__global__ void __launch_bounds__(E_WARPS * 32, EP_BLOCKS) k_estep_pair_ws(int* __restrict__ cnt, const PairDesc* __restrict__ pdesc,
                                                                            const double* __restrict__ phil2 /*[6][cap] by pair position*/,
                                                                            int64_t cap, RegionTables rt, int64_t NI, double* __restrict__ ES) {
    extern __shared__ __align__(128) unsigned char ep_smem[];
    PairTile* tiles = reinterpret_cast<PairTile*>(ep_smem);
    EmisRing* rings = reinterpret_cast<EmisRing*>(ep_smem + (size_t)E_WARPS * sizeof(PairTile));
    PairStage* stages = reinterpret_cast<PairStage*>(ep_smem + (size_t)E_WARPS * (sizeof(PairTile) + sizeof(EmisRing)));
    unsigned long long* mbars = reinterpret_cast<unsigned long long*>(ep_smem + (size_t)E_WARPS * (sizeof(PairTile) + sizeof(EmisRing) + sizeof(PairStage)));
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, h = lane >> 4, i = lane & 15;
    PairStage& stg = stages[w];
    EmisSrc<false> es;
    es.mbar = (unsigned)__cvta_generic_to_shared(&mbars[4 * w]);
    es.ring = (unsigned)__cvta_generic_to_shared(&rings[w].r[0][0][0]);
    es.phase = 0u;
    if (lane == 0) {
        mbar_init(es.mbar, 1u); mbar_init(es.mbar + 8u, 1u); mbar_init(es.mbar + 16u, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int n_pos = __ldg(cnt), n_pairs = n_pos + __ldg(cnt + 1);
    auto take = [&]() -> int {
        int q = 0;
        if (lane == 0) q = atomicAdd(cnt + 2, 1);
        return __shfl_sync(0xffffffffu, q, 0);
    };
    auto position = [&](int pq) -> int { return pq < n_pos ? pq : (int)cap - 1 - (pq - n_pos); };
    int pq = take();
    if (pq >= n_pairs) return;
    // the first pair of this warp: nothing to overlap with
    stage_desc(stg, pdesc, phil2, cap, position(pq), lane);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    stage_tables(stg, rt, reinterpret_cast<const longlong2*>(&stg.d[1])->x, stg.d[0].z, lane);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
#pragma unroll 1
    while (true) {
        const int4 d0 = stg.d[0];
        const longlong2 d1 = *reinterpret_cast<const longlong2*>(&stg.d[1]);
        const double a = stg.phi[3 * h], b = stg.phi[3 * h + 1], c = stg.phi[3 * h + 2];
        const int L = d0.z, M = d0.w & ITEM_M_MASK, mode = (d0.w >> ITEM_MODE_SHIFT) & 3;
        const int64_t g0 = d1.x;
        const bool bpos = __shfl_sync(0xffffffffu, (int)(__double2hiint(c) >= 0), 0) != 0;   // also orders the reads above before the refill
        const double cabs = fabs(c);
        const double* ew_r = rt.ew + d1.y;
        const int nq = take();
        const bool has_next = nq < n_pairs;
        if (has_next) stage_desc(stg, pdesc, phil2, cap, position(nq), lane);
        double acc0 = 0, acc1 = 0, acc2 = 0;
        const int nchunk = (M + 31) >> 5;
        for (int ch = 0; ch < nchunk; ++ch) {
            double o0[3], o1[3];
            // the next pair's tables replace the staged ones after the LAST chunk's first tile; earlier chunks leave them in place
            const bool last = ch == nchunk - 1;
            if (mode != 0) {
                const size_t eo = (size_t)ch * L * 32;
                es.p = ew_r + eo;
                es.pi = rt.ewi + d1.y + eo;
                estep_pair_chunk_ws(tiles[w], rings[w], stg, rt, L, lane, rt.na + g0, rt.nb + g0, rt.invd + g0, es, a, b, cabs, bpos, mode - 1,
                                    true, has_next && last, o0, o1);
            } else {
                const double* p0 = ew_r + (size_t)ch * L * 32 + i;
                if (bpos) estep_pair_slow<true>(tiles[w], L, lane, rt.na + g0, rt.nb + g0, rt.invd + g0, p0, p0 + 16, a, b, cabs, o0, o1);
                else estep_pair_slow<false>(tiles[w], L, lane, rt.na + g0, rt.nb + g0, rt.invd + g0, p0, p0 + 16, a, b, cabs, o0, o1);
                if (has_next && last) {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncwarp();
                    stage_tables(stg, rt, reinterpret_cast<const longlong2*>(&stg.d[1])->x, stg.d[0].z, lane);
                }
            }
            const bool live0 = ch * 32 + i < M, live1 = ch * 32 + 16 + i < M;
            double v0 = (live0 ? o0[0] : 0.0) + (live1 ? o1[0] : 0.0);
            double v1 = (live0 ? o0[1] : 0.0) + (live1 ? o1[1] : 0.0);
            double v2 = (live0 ? o0[2] : 0.0) + (live1 ? o1[2] : 0.0);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                v0 += __shfl_xor_sync(0xffffffffu, v0, o);
                v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                v2 += __shfl_xor_sync(0xffffffffu, v2, o);
            }
            acc0 += v0; acc1 += v1; acc2 += v2;
        }
        const int inst = h ? d0.y : d0.x;
        pair_poison(a, b, c, acc0, acc1, acc2);
        if (i == 0 && inst >= 0) { ES[inst] = acc0; ES[NI + inst] = acc1; ES[2 * NI + inst] = acc2; }
        if (!has_next) break;
        asm volatile("cp.async.wait_group 0;" ::: "memory");      // the next pair's tables (group B)
        __syncwarp();
    }
}

// Pairs of the active list (see above).  Thread per list position; leaders (even rank within their region's run and sign class)
// append one pair each through a warp-aggregated counter — the order of the pair list follows the active list only roughly,
// which changes scheduling, never results.
__global__ void k_make_pairs(int n_items, const int* __restrict__ items, int n_init, const RegionDesc* __restrict__ rdesc,
                             const double* __restrict__ phi /*[3][NI]*/, int64_t NI, PairDesc* __restrict__ pdesc, double* __restrict__ phil2,
                             int64_t cap, const int* __restrict__ n_dev, int* __restrict__ n_pairs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_dev) n_items = __ldg(n_dev);
    bool leader = false;
    int inst = 0, partner = -1, u = 0;
    double c = 0.0;
    if (i < n_items) {
        inst = items ? items[i] : i;
        u = inst / n_init;
        c = phi[2 * NI + inst];
        const bool cneg = __double2hiint(c) < 0, aneg = __double2hiint(phi[inst]) < 0;
        // Pairing order within the region's run and the sign class of c: starts with a ≥ 0 in list order, then starts with a < 0
        // — two starts that agree on the sign of a mostly agree on the sign of α_l = N_l (a + b ρ_l), i.e. on the block body.
        int pos_before = 0, neg_before = 0, pos_after = 0;
        int next_pos = -1, next_neg = -1;                  // first classmates after this one
        for (int j = 1; j < n_init && i - j >= 0; ++j) {
            const int o = items ? items[i - j] : i - j;
            if (o / n_init != u) break;
            if ((__double2hiint(phi[2 * NI + o]) < 0) != cneg) continue;
            if (__double2hiint(phi[o]) < 0) ++neg_before; else ++pos_before;
        }
        for (int j = 1; j < n_init && i + j < n_items; ++j) {
            const int o = items ? items[i + j] : i + j;
            if (o / n_init != u) break;
            if ((__double2hiint(phi[2 * NI + o]) < 0) != cneg) continue;
            if (__double2hiint(phi[o]) < 0) { if (next_neg < 0) next_neg = o; }
            else { ++pos_after; if (next_pos < 0) next_pos = o; }
        }
        // first a < 0 classmate of the whole run (needed by the last a ≥ 0 start when their number is odd)
        int first_neg = next_neg;
        if (neg_before > 0) {
            for (int j = 1; j < n_init && i - j >= 0; ++j) {
                const int o = items ? items[i - j] : i - j;
                if (o / n_init != u) break;
                if ((__double2hiint(phi[2 * NI + o]) < 0) == cneg && __double2hiint(phi[o]) < 0) first_neg = o;
            }
        }
        const int rank = aneg ? pos_before + pos_after + neg_before : pos_before;
        leader = (rank & 1) == 0;
        if (leader) partner = aneg ? next_neg : (next_pos >= 0 ? next_pos : first_neg);
    }
    const bool lneg = leader && __double2hiint(c) < 0;
    const unsigned mask_p = __ballot_sync(0xffffffffu, leader && !lneg), mask_n = __ballot_sync(0xffffffffu, lneg);
    if ((mask_p | mask_n) == 0u) return;
    const int lane = threadIdx.x & 31;
    int base_p = 0, base_n = 0;
    if (mask_p) {
        const int first = __ffs(mask_p) - 1;
        if (lane == first) base_p = atomicAdd(n_pairs, __popc(mask_p));
        base_p = __shfl_sync(0xffffffffu, base_p, first);
    }
    if (mask_n) {
        const int first = __ffs(mask_n) - 1;
        if (lane == first) base_n = atomicAdd(n_pairs + 1, __popc(mask_n));
        base_n = __shfl_sync(0xffffffffu, base_n, first);
    }
    if (!leader) return;
    const unsigned below = (1u << lane) - 1u;
    // c ≥ 0: k-th pair at position k; c < 0: k-th pair at position cap − 1 − k (k_estep_pair reads them in the same order)
    const int p = lneg ? (int)cap - 1 - (base_n + __popc(mask_n & below)) : base_p + __popc(mask_p & below);
    const RegionDesc rd = rdesc[u];
    const int ib = partner >= 0 ? partner : inst;
    const double cB = phi[2 * NI + ib];
    // rescale policy: the more careful of the two instances' (see k_make_desc); NaN bounds or a NaN c choose the slow path
    const double B = (double)rd.llrmax + 2.0 * fmax(fabs(c), fabs(cB)) * (double)rd.idmax;
    int mode = (B <= (double)E_BOUND_2) ? 2 : ((B <= (double)E_BOUND_1) ? 1 : 0);
    if (c != c || cB != cB) mode = 0;
    PairDesc d;
    d.instA = inst; d.instB = partner; d.L = rd.L; d.M = rd.M | (mode << ITEM_MODE_SHIFT); d.g0 = rd.g0; d.ew_off = rd.ew_off;
    d.ridx = rd.ridx; d.pad0 = d.pad1 = d.pad2 = 0;
    pdesc[p] = d;
    phil2[p] = phi[inst]; phil2[cap + p] = phi[NI + inst]; phil2[2 * cap + p] = c;
    phil2[3 * cap + p] = phi[ib]; phil2[4 * cap + p] = phi[NI + ib]; phil2[5 * cap + p] = cB;
}

// ---------------------------------------------------------------------------------------------------
// score: Q = (1/M) Σ_m log p(y_m; ϕ) = (1/M) Σ_m [base_m + log Zc_m,rel] − log Z_rel
// (the |α|,|β| offsets of the relative potentials are identical in Zc and Z and cancel)
// ---------------------------------------------------------------------------------------------------
template <bool BPOS>
__device__ __forceinline__ double logz_chunk(SiteTile& st, int L, int lane, const double* __restrict__ na_g,
                                             const double* __restrict__ nb_g, const double* __restrict__ invd_g,
                                             const double* __restrict__ p /*r of this lane's read, row stride EW_ROW; null → r = 1*/, double a,
                                             double b, double cabs) {
    double fp = 1.0, fm = 1.0, sc = 1.0;           // all-ones message; group 0 carries t = 0 (see fill_tile)
    int esum = 0, e;
    for (int l0 = 0; l0 < L; l0 += TILE) {
        int n = min(TILE, L - l0);
        __syncwarp();
        fill_tile(st, l0, n, lane, na_g, nb_g, invd_g, a, b, cabs, BPOS);
        __syncwarp();
        // emission weights of 8 groups are loaded before they are used: the chain step is a handful of dependent FP64
        // instructions, so a load per step sat on the critical path (same arithmetic, same order: results unchanged)
        for (int j0 = 0; j0 < n; j0 += 8) {
            double rr[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) rr[k] = (p && j0 + k < n) ? __ldg(p + (size_t)(l0 + j0 + k) * EW_ROW) : 1.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int j = j0 + k;
                if (j < n) {
                    const double2 w = st.w[j];
                    double t = st.tn[j].x;
                    double phiP = (w.y * rr[k]) * sc, phiM = w.x * sc;
                    double u = BPOS ? fp : fm, v = BPOS ? fm : fp;
                    fp = phiP * fma(t, v, u);
                    fm = phiM * fma(t, u, v);
                    sc = cpn_inv_pow2_of(fp + fm, &e); esum += e;
                }
            }
        }
    }
    return log((fp + fm) * sc) + (double)esum * 0.6931471805599453;
}

// item → (instance); computes Q[inst] for instances whose flag[inst] == want (or all if flag == null)
__global__ void __launch_bounds__(E_WARPS * 32) k_score(int n_items, const int* __restrict__ items, int n_init, RegionTables rt,
                                                         const double* __restrict__ phi, int64_t NI, const int* __restrict__ flag,
                                                         int want, double* __restrict__ Q) {
    __shared__ SiteTile tiles[E_WARPS];
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int item = blockIdx.x * E_WARPS + w;
    if (item >= n_items) return;
    int inst = items ? items[item] : item;
    if (flag && flag[inst] != want) {
        if (lane == 0) Q[inst] = -cpn_inf();
        return;
    }
    const RegionDesc rd = rt.rdesc[inst / n_init];
    const int64_t g0 = rd.g0;
    const int L = rd.L;
    const int64_t M = rd.M;
    double a = phi[inst], b = phi[NI + inst], c = phi[2 * NI + inst];
    bool bpos = c >= 0.0;
    double cabs = fabs(c);
    const double* ew_r = rt.ew + rd.ew_off;
    const double* rb_r = rt.rb + rd.rb_off;
    int nchunk = (int)((M + 31) >> 5);
    double acc = 0.0;
    // The data-free chain log Z(ϕ) rides in lane 31 of the last chunk whenever that lane has no read (M not a multiple of 32):
    // same instructions as a read with all emission weights 1, so the separate pass (a whole warp computing one chain) goes.
    const bool spare = (M & 31) != 0;
    double lz0 = 0.0;
    for (int ch = 0; ch < nchunk; ++ch) {
        int q = ch * 32 + lane;
        const bool live = (int64_t)ch * 32 + lane < M;
        if (rd.ridx >= 0) q = __ldg(rt.ridx + ((size_t)(rd.ridx + ch) << 5) + lane);
        const double* p = ew_r + (size_t)(q >> 5) * L * EW_ROW + (q & 31);
        if (spare && ch == nchunk - 1 && lane == 31) p = nullptr;
        double lz = bpos ? logz_chunk<true>(tiles[w], L, lane, rt.na + g0, rt.nb + g0, rt.invd + g0, p, a, b, cabs)
                         : logz_chunk<false>(tiles[w], L, lane, rt.na + g0, rt.nb + g0, rt.invd + g0, p, a, b, cabs);
        acc += cpn_warp_sum(live ? rb_r[q] + lz : 0.0);
        if (spare && ch == nchunk - 1) lz0 = __shfl_sync(0xffffffffu, lz, 31);
    }
    if (!spare)      // every lane of the last chunk holds a read: the marginal chain needs its own pass (all lanes compute the same value)
        lz0 = bpos ? logz_chunk<true>(tiles[w], L, lane, rt.na + g0, rt.nb + g0, rt.invd + g0, nullptr, a, b, cabs)
                   : logz_chunk<false>(tiles[w], L, lane, rt.na + g0, rt.nb + g0, rt.invd + g0, nullptr, a, b, cabs);
    if (lane == 0) Q[inst] = acc / (double)M - lz0;
}

// ---------------------------------------------------------------------------------------------------
// M-step
// ---------------------------------------------------------------------------------------------------
// log Z_rel(ϕ) (value only) for the FD mode: serial sweep on one lane.
__device__ double logz_serial(const double* __restrict__ na_g, const double* __restrict__ nb_g, const double* __restrict__ invd_g,
                              int L, double a, double b, double c) {
    // offsets Σ|α_l| + Σ|β_l| are NOT constant in ϕ here, so the FD mode needs the full log Z.
    bool bpos = c >= 0.0;
    double cabs = fabs(c);
    double fp, fm, sc, off = 0.0;
    int esum = 0, e;
    {
        double alpha = fma(b, nb_g[0], a * na_g[0]);
        double ex = exp(-2.0 * fabs(alpha));
        off += fabs(alpha);
        fp = alpha >= 0 ? 1.0 : ex; fm = alpha >= 0 ? ex : 1.0;
        sc = cpn_inv_pow2_of(fp + fm, &e); esum += e;
    }
    for (int l = 1; l < L; ++l) {
        double alpha = fma(b, nb_g[l], a * na_g[l]);
        double ex = exp(-2.0 * fabs(alpha));
        double beta = cabs * invd_g[l - 1];
        double t = exp(-2.0 * beta);
        off += fabs(alpha) + beta;
        double wp = (alpha >= 0 ? 1.0 : ex) * sc, wm = (alpha >= 0 ? ex : 1.0) * sc;
        double u = bpos ? fp : fm, v = bpos ? fm : fp;
        fp = wp * fma(t, v, u);
        fm = wm * fma(t, u, v);
        sc = cpn_inv_pow2_of(fp + fm, &e); esum += e;
    }
    return log((fp + fm) * sc) + (double)esum * 0.6931471805599453 + off;
}

// Gradient g[3] and Hessian H (aa,ab,ac,bb,bc,cc) of log Z(ϕ) by one forward sweep with first- and
// second-order tangents.  c<0 is mapped onto c>0 by the gauge x_l → (−1)^l x_l (flips the sign of α_l,
// N_l, N_lρ_l on odd sites and of 1/d_l everywhere), so the recursion always has ψ(aligned)=1, ψ(opposite)=t.
struct SitePot { double na, nb, invd, invd2, na2, nanb, nb2, w0, w1, t; };

__device__ __forceinline__ void prefetch_site(const SweepSite* p) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));              // 32-byte record, 32-byte aligned: one sector
}
__device__ __forceinline__ SweepSite load_site(const SweepSite* __restrict__ p) {
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 x0 = __ldg(q), x1 = __ldg(q + 1);
    SweepSite r;
    r.na = x0.x; r.nb = x0.y; r.id = x1.x; r.id2 = x1.y;
    return r;
}
#ifndef MS_EXP_CLAMPED
#define MS_EXP_CLAMPED 1
#endif
#if MS_EXP_CLAMPED
#define MS_EXP cpn_exp_neg_clamped          // a non-finite trial point never reaches the caller's results: see `stopped` in the solvers
#else
#define MS_EXP cpn_exp_neg
#endif
// Potentials of group l at ϕ (gauge for c<0: sign flip of N_l, N_lρ_l on odd groups and of 1/d everywhere).
__device__ __forceinline__ SitePot site_pot(const SweepSite& r, int l, bool neg, double a, double b, double cabs) {
    const int flipbit = (neg && (l & 1)) ? (int)0x80000000 : 0;
    const int negbit = neg ? (int)0x80000000 : 0;
    SitePot p;
    p.na = __hiloint2double(__double2hiint(r.na) ^ flipbit, __double2loint(r.na));
    p.nb = __hiloint2double(__double2hiint(r.nb) ^ flipbit, __double2loint(r.nb));
    p.invd = __hiloint2double(__double2hiint(r.id) ^ negbit, __double2loint(r.id));
    p.invd2 = r.id2; p.na2 = r.na * r.na; p.nanb = r.na * r.nb; p.nb2 = r.nb * r.nb;
    double alpha = fma(b, p.nb, a * p.na);
    double ex = MS_EXP(-2.0 * fabs(alpha));
    const bool pos = __double2hiint(alpha) >= 0;
    p.w0 = pos ? ex : 1.0; p.w1 = pos ? 1.0 : ex;
    p.t = MS_EXP(-2.0 * (cabs * r.id));
    return p;
}

// The loop is software-pipelined by hand: while the recursion consumes group l, the record of group l+2 is being
// loaded and the two exponentials of group l+1 are evaluated (they do not depend on the message), so the only serial
// chain is the 2-state recursion itself.
// RING = true (work-stealing kernel): the records travel global → shared memory with cp.async into a private 4-slot ring
// per lane, issued three groups ahead.  An asynchronous copy has no destination register, so the compiler cannot sink it
// next to its use the way it does with an ordinary load whose value is only needed one iteration later.
constexpr int RING_SLOTS = 4;
__device__ __forceinline__ void ring_issue(double2* ring, int slot, const SweepSite* src) {
    // two 16-byte halves of the 32-byte record → ring[(slot*2 + half) * M_THREADS + tid]
    unsigned d0 = (unsigned)__cvta_generic_to_shared(ring + (slot * 2 + 0) * M_THREADS + threadIdx.x);
    unsigned d1 = (unsigned)__cvta_generic_to_shared(ring + (slot * 2 + 1) * M_THREADS + threadIdx.x);
    const char* g = reinterpret_cast<const char*>(src);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d0), "l"(g));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d1), "l"(g + 16));
    asm volatile("cp.async.commit_group;");
}
__device__ __forceinline__ SweepSite ring_read(const double2* ring, int slot) {
    double2 x0 = ring[(slot * 2 + 0) * M_THREADS + threadIdx.x], x1 = ring[(slot * 2 + 1) * M_THREADS + threadIdx.x];
    SweepSite r;
    r.na = x0.x; r.nb = x0.y; r.id = x1.x; r.id2 = x1.y;
    return r;
}

template <bool RING>
__device__ void eval_grad_hess(const SweepSite* __restrict__ tbl, int L, double a, double b, double c, double g[3], double H[6],
                               double2* ring = nullptr) {
    const bool neg = __double2hiint(c) < 0;
    const double cabs = fabs(c);
    double f[2], ga[2], gb[2], gc[2], haa[2], hab[2], hac[2], hbb[2], hbc[2], hcc[2];
    double sc;
    SweepSite raw;
    if (RING) {
        // prologue: records 1, 2, 3 in flight (slots 1, 2, 3); record 0 is read directly
        ring_issue(ring, 1, tbl + (L > 1 ? 1 : 0));
        ring_issue(ring, 2, tbl + (L > 2 ? 2 : 0));
        ring_issue(ring, 3, tbl + (L > 3 ? 3 : 0));
    } else {
        raw = load_site(tbl + (L > 1 ? 1 : 0));
    }
    {
        SitePot p0 = site_pot(load_site(tbl), 0, neg, a, b, cabs);
        double w[2] = {p0.w0, p0.w1};
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            double sg = x ? 1.0 : -1.0;
            f[x] = w[x];
            ga[x] = sg * p0.na * w[x]; gb[x] = sg * p0.nb * w[x]; gc[x] = 0.0;
            haa[x] = p0.na2 * w[x]; hab[x] = p0.nanb * w[x]; hbb[x] = p0.nb2 * w[x];
            hac[x] = 0.0; hbc[x] = 0.0; hcc[x] = 0.0;
        }
        sc = cpn_inv_pow2_of(f[0] + f[1]);
    }
    if (RING) {
        asm volatile("cp.async.wait_group 2;");                     // record 1 has landed
        raw = ring_read(ring, 1);
    }
    SitePot pot = site_pot(raw, 1, neg, a, b, cabs);
    if (!RING) raw = load_site(tbl + (L > 2 ? 2 : 0));
    // unrolled by two: the 20 doubles of loop-carried state ping-pong between two register sets instead of being copied back
    // (34 register moves per group otherwise, each an issue slot next to the 2-cycle FP64 issues)
#pragma unroll 2
    for (int l = 1; l < L; ++l) {
        const SitePot p = pot;
        SweepSite rnext;
        if (RING) {
            // in flight during this trip: records l+1, l+2, l+3 (the new one goes to the slot record l-1 vacated);
            // waiting until at most two groups are pending guarantees record l+1 has landed
            ring_issue(ring, (l + 3) & (RING_SLOTS - 1), tbl + (l + 3 < L ? l + 3 : 0));
            asm volatile("cp.async.wait_group 2;");
            rnext = ring_read(ring, (l + 1) & (RING_SLOTS - 1));
        } else {
            rnext = raw;
            prefetch_site(tbl + (l + 4 < L ? l + 4 : 0));           // pull the record of group l+4 into L1 (no register cost)
            raw = load_site(tbl + (l + 2 < L ? l + 2 : 0));        // consumed by site_pot() in the next iteration
        }
        pot = site_pot(rnext, l + 1, neg, a, b, cabs);              // independent of the recursion below
        const double t = p.t, invd = p.invd, invd2 = p.invd2, invd_2 = p.invd + p.invd;
        double w[2] = {p.w0 * sc, p.w1 * sc};
        double nf[2], nga[2], ngb[2], ngc[2], nhaa[2], nhab[2], nhac[2], nhbb[2], nhbc[2], nhcc[2];
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            const int s = x, o = 1 - x;
            const double sna = x ? p.na : -p.na, snb = x ? p.nb : -p.nb;
            double S0 = fma(t, f[o], f[s]);
            double Sa = fma(t, ga[o], ga[s]);
            double Sb = fma(t, gb[o], gb[s]);
            double gcl = fma(t, gc[o], gc[s]);
            double D0 = fma(2.0, f[s], -S0);
            double Sc = fma(invd, D0, gcl);
            double Saa = fma(t, haa[o], haa[s]);
            double Sab = fma(t, hab[o], hab[s]);
            double Sbb = fma(t, hbb[o], hbb[s]);
            double Sac = fma(invd, fma(2.0, ga[s], -Sa), fma(t, hac[o], hac[s]));
            double Sbc = fma(invd, fma(2.0, gb[s], -Sb), fma(t, hbc[o], hbc[s]));
            double Scc = fma(invd_2, fma(2.0, gc[s], -gcl), fma(invd2, S0, fma(t, hcc[o], hcc[s])));
            double ph = w[x];
            nf[x] = ph * S0;
            nga[x] = ph * fma(sna, S0, Sa);
            ngb[x] = ph * fma(snb, S0, Sb);
            ngc[x] = ph * Sc;
            nhaa[x] = ph * fma(p.na2, S0, fma(sna, Sa + Sa, Saa));
            nhab[x] = ph * fma(p.nanb, S0, fma(sna, Sb, fma(snb, Sa, Sab)));
            nhbb[x] = ph * fma(p.nb2, S0, fma(snb, Sb + Sb, Sbb));
            nhac[x] = ph * fma(sna, Sc, Sac);
            nhbc[x] = ph * fma(snb, Sc, Sbc);
            nhcc[x] = ph * Scc;
        }
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            f[x] = nf[x]; ga[x] = nga[x]; gb[x] = ngb[x]; gc[x] = ngc[x];
            haa[x] = nhaa[x]; hab[x] = nhab[x]; hac[x] = nhac[x]; hbb[x] = nhbb[x]; hbc[x] = nhbc[x]; hcc[x] = nhcc[x];
        }
        sc = cpn_inv_pow2_of(f[0] + f[1]);
    }
    double inv = 1.0 / (f[0] + f[1]);
    g[0] = (ga[0] + ga[1]) * inv; g[1] = (gb[0] + gb[1]) * inv; g[2] = (gc[0] + gc[1]) * inv;
    H[0] = fma(-g[0], g[0], (haa[0] + haa[1]) * inv);
    H[1] = fma(-g[0], g[1], (hab[0] + hab[1]) * inv);
    H[2] = fma(-g[0], g[2], (hac[0] + hac[1]) * inv);
    H[3] = fma(-g[1], g[1], (hbb[0] + hbb[1]) * inv);
    H[4] = fma(-g[1], g[2], (hbc[0] + hbc[1]) * inv);
    H[5] = fma(-g[2], g[2], (hcc[0] + hcc[1]) * inv);
    if (RING) asm volatile("cp.async.wait_group 0;");               // drain before the ring is reused by the next sweep
}

struct MProblem {
    const double *na, *nb, *invd;
    const SweepSite* sq;
    int L;
    double t0, t1, t2;     // ES/M
    int mode;
    int evals;             // log Z-equivalent evaluations (FD) / derivative sweeps (analytic)
};

// residual U(ϕ) = ∇logZ(ϕ) − ES/M  (em_algorithm.jl:152-157); FD mode: Calculus.gradient central rule.
__device__ void residual_fd(MProblem& pb, double* x, double* U) {
    for (int i = 0; i < 3; ++i) {
        double epsilon = CBRT_EPS * fmax(1.0, fabs(x[i]));
        double oldx = x[i];
        x[i] = oldx + epsilon;
        double fp = logz_serial(pb.na, pb.nb, pb.invd, pb.L, x[0], x[1], x[2]);
        x[i] = oldx - epsilon;
        double fm = logz_serial(pb.na, pb.nb, pb.invd, pb.L, x[0], x[1], x[2]);
        x[i] = oldx;
        U[i] = (fp - fm) / (epsilon + epsilon);
    }
    pb.evals += 6;
    U[0] -= pb.t0; U[1] -= pb.t1; U[2] -= pb.t2;
}
// FiniteDiff central Jacobian of the residual: ε_k = max(cbrt(eps)|x_k|, cbrt(eps))
__device__ void jacobian_fd(MProblem& pb, double* x, double J[3][3]) {
    double x1[3] = {x[0], x[1], x[2]}, f1[3], f0[3];
    for (int k = 0; k < 3; ++k) {
        double x_save = x[k];
        double epsilon = fmax(CBRT_EPS * fabs(x_save), CBRT_EPS);
        x1[k] = x_save + epsilon; residual_fd(pb, x1, f1);
        x[k] = x_save - epsilon;  residual_fd(pb, x, f0);
        for (int i = 0; i < 3; ++i) J[i][k] = (f1[i] - f0[i]) / (2 * epsilon);
        x1[k] = x_save; x[k] = x_save;
    }
}
__device__ void value_jacobian(MProblem& pb, double* x, double* F, double J[3][3]) {
    if (pb.mode == CPN_MSTEP_ANALYTIC) {
        double g[3], H[6];
        eval_grad_hess<false>(pb.sq, pb.L, x[0], x[1], x[2], g, H);
        pb.evals += 1;
        F[0] = g[0] - pb.t0; F[1] = g[1] - pb.t1; F[2] = g[2] - pb.t2;
        J[0][0] = H[0]; J[0][1] = H[1]; J[0][2] = H[2];
        J[1][0] = H[1]; J[1][1] = H[3]; J[1][2] = H[4];
        J[2][0] = H[2]; J[2][1] = H[4]; J[2][2] = H[5];
    } else {
        residual_fd(pb, x, F);
        jacobian_fd(pb, x, J);
    }
}

__device__ __forceinline__ double wdot3(const double* wx, const double* x, const double* wy, const double* y) {
    double out = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) out += wx[i] * x[i] * wy[i] * y[i];
    return out;
}
__device__ __forceinline__ double wnorm3(const double* w, const double* x) { return sqrt(wdot3(w, x, w, x)); }
__device__ __forceinline__ double sumabs2_3(const double* x) { return x[0] * x[0] + x[1] * x[1] + x[2] * x[2]; }
__device__ __forceinline__ double norminf3(const double* x) { return fmax(fabs(x[0]), fmax(fabs(x[1]), fabs(x[2]))); }
__device__ __forceinline__ bool anynan3(const double* x) { return x[0] != x[0] || x[1] != x[1] || x[2] != x[2]; }
__device__ __forceinline__ bool allfinite3(const double* x) { return isfinite(x[0]) && isfinite(x[1]) && isfinite(x[2]); }

// J \ r, LU with partial pivoting; false on an exactly singular pivot.
__device__ bool lu_solve3(const double Jin[3][3], const double* r, double* out) {
    double A[3][3], b[3] = {r[0], r[1], r[2]};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) A[i][j] = Jin[i][j];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int piv = k; double best = fabs(A[k][k]);
#pragma unroll
        for (int i = k + 1; i < 3; ++i) if (fabs(A[i][k]) > best) { best = fabs(A[i][k]); piv = i; }
        if (!(best > 0.0) && !(best != best)) return false;
#pragma unroll
        for (int i = k + 1; i < 3; ++i) {
            if (piv == i) {
#pragma unroll
                for (int j = 0; j < 3; ++j) { double tmp = A[k][j]; A[k][j] = A[i][j]; A[i][j] = tmp; }
                double tb = b[k]; b[k] = b[i]; b[i] = tb;
            }
        }
#pragma unroll
        for (int i = k + 1; i < 3; ++i) {
            double m = A[i][k] / A[k][k];
#pragma unroll
            for (int j = k + 1; j < 3; ++j) A[i][j] -= m * A[k][j];
            b[i] -= m * b[k];
        }
    }
    out[2] = b[2] / A[2][2];
    out[1] = (b[1] - A[1][2] * out[2]) / A[1][1];
    out[0] = (b[0] - A[0][1] * out[1] - A[0][2] * out[2]) / A[0][0];
    return true;
}
__device__ __forceinline__ void matvec3(const double J[3][3], const double* v, double* out) {
#pragma unroll
    for (int i = 0; i < 3; ++i) out[i] = J[i][0] * v[0] + J[i][1] * v[1] + J[i][2] * v[2];
}
// NLsolve dogleg! (trust_region.jl).  An exactly singular J (measure zero) takes the Cauchy branch.
__device__ void dogleg3(double* p, const double* r, const double* d, const double J[3][3], double delta) {
    double p_i[3], p_c[3];
    bool ok = lu_solve3(J, r, p_i);
    if (ok) {
#pragma unroll
        for (int i = 0; i < 3; ++i) p_i[i] = -p_i[i];
        if (wnorm3(d, p_i) <= delta) { p[0] = p_i[0]; p[1] = p_i[1]; p[2] = p_i[2]; return; }
    }
    double g[3], Jg[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) g[j] = (J[0][j] * r[0] + J[1][j] * r[1] + J[2][j] * r[2]) / (d[j] * d[j]);
    matvec3(J, g, Jg);
    double wg = wnorm3(d, g);
    double coef = -(wg * wg) / sumabs2_3(Jg);
#pragma unroll
    for (int i = 0; i < 3; ++i) p_c[i] = coef * g[i];
    if (!ok || wnorm3(d, p_c) >= delta) {
        double s = -delta / wnorm3(d, g);
#pragma unroll
        for (int i = 0; i < 3; ++i) p[i] = g[i] * s;
    } else {
        double p_diff[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) p_diff[i] = p_i[i] - p_c[i];
        double b = 2 * wdot3(d, p_c, d, p_diff);
        double wpd = wnorm3(d, p_diff), a = wpd * wpd, wpc = wnorm3(d, p_c);
        double tau = (-b + sqrt(b * b - 4 * a * (wpc * wpc - delta * delta))) / (2 * a);
#pragma unroll
        for (int i = 0; i < 3; ++i) p[i] = p_c[i] + tau * p_diff[i];
    }
}
// NLsolve.nlsolve(f!, x0; iterations=20, ftol=1e-3): trust region, autoscale, factor 1, xtol 0.
// Returns converged(sol); *threw mirrors the IsFiniteException on a non-finite initial residual.
__device__ bool trust_region_solve(MProblem& pb, const double* x0, double* x, int* n_iters, bool* threw) {
    const double ftol = 1e-3, eta = 1e-4;
    const int iterations = 20;
    double r[3], J[3][3], fval[3], d[3], p[3], xold[3];
    *threw = false;
    x[0] = x0[0]; x[1] = x0[1]; x[2] = x0[2];
    value_jacobian(pb, x, fval, J);
    r[0] = fval[0]; r[1] = fval[1]; r[2] = fval[2];
    *n_iters = 0;
    if (!allfinite3(r)) { *threw = true; return false; }
    bool x_conv = false, f_conv = norminf3(fval) <= ftol;
    bool stopped = !allfinite3(x) || anynan3(fval);      // NLsolve stops on NaN; a non-finite x gives NaN at its next evaluation
    bool conv = f_conv;
    if (conv) return true;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        d[j] = sqrt(J[0][j] * J[0][j] + J[1][j] * J[1][j] + J[2][j] * J[2][j]);
        if (d[j] == 0.0) d[j] = 1.0;
    }
    double delta = wnorm3(d, x);
    if (delta == 0.0) delta = 1.0;
    int it = 0;
    while (!stopped && !conv && it < iterations) {
        ++it;
        dogleg3(p, r, d, J, delta);
#pragma unroll
        for (int i = 0; i < 3; ++i) { xold[i] = x[i]; x[i] += p[i]; }
        double Jn[3][3];
        if (pb.mode == CPN_MSTEP_ANALYTIC) value_jacobian(pb, x, fval, Jn);   // F and J from the same sweep
        else residual_fd(pb, x, fval);
        double rp[3];
        matvec3(J, p, rp);
#pragma unroll
        for (int i = 0; i < 3; ++i) rp[i] += r[i];
        double rho = (sumabs2_3(r) - sumabs2_3(fval)) / (sumabs2_3(r) - sumabs2_3(rp));
        if (rho > eta) {
            r[0] = fval[0]; r[1] = fval[1]; r[2] = fval[2];
            if (pb.mode == CPN_MSTEP_ANALYTIC) {
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) J[i][j] = Jn[i][j];
            } else {
                double Fd[3];
                value_jacobian(pb, x, Fd, J);
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                double nj = sqrt(J[0][j] * J[0][j] + J[1][j] * J[1][j] + J[2][j] * J[2][j]);
                d[j] = fmax(0.1 * d[j], nj);
            }
            double dx[3] = {x[0] - xold[0], x[1] - xold[1], x[2] - xold[2]};
            x_conv = norminf3(dx) <= 0.0;
            f_conv = norminf3(r) <= ftol;
            conv = x_conv || f_conv;
        } else {
#pragma unroll
            for (int i = 0; i < 3; ++i) x[i] -= p[i];
            x_conv = false; conv = false;
        }
        if (rho < 0.1) delta = delta / 2;
        else if (rho >= 0.9) delta = 2 * wnorm3(d, p);
        else if (rho >= 0.5) delta = fmax(delta, 2 * wnorm3(d, p));
        stopped = !allfinite3(x) || anynan3(fval);      // NLsolve stops on NaN; a non-finite x gives NaN at its next evaluation
    }
    *n_iters = it;
    return x_conv || f_conv;
}

struct EmState {
    double* phi;        // [3][NI] current iterate
    double* ES;         // [3][NI]
    int* iter;          // [NI] the reference's counter i (starts at 1)
    int* status;        // [NI] 0 active, 1 converged, 2 not converged
    int* cnt_em;        // [NI] EM iterations done
    int* cnt_sol;       // [NI] solver iterations
    int* cnt_eval;      // [NI] log Z evaluations / derivative sweeps
    int64_t NI;
    double* gh;                 // [9][NI] ∇logZ (3) and its Hessian (6) at the instance's CURRENT ϕ, when gh_valid
    unsigned char* gh_valid;    // [NI]
    unsigned char* numerr;      // [NI] 1 once an M-step of the instance met a non-finite residual or iterate (the reference
                                //      catches the exception and keeps ϕ, em_algorithm.jl:160-172); feeds CPN_FIT_NUMERICAL
};

// One em_iter M-step + the em_inst bookkeeping (em_algorithm.jl:152-183, :317-341) for every active
// instance; survivors are appended (block-ordered) to next_items.
__global__ void __launch_bounds__(M_THREADS) k_mstep(int n_items, const int* __restrict__ items, int n_init, RegionTables rt,
                                                      EmState st, int max_em_iters, double amax, double bmax, double cmax, int mode,
                                                      int* __restrict__ next_items, int* __restrict__ n_next,
                                                      const int* __restrict__ n_dev) {
    int item = blockIdx.x * M_THREADS + threadIdx.x;
    if (n_dev) n_items = __ldg(n_dev);
    bool survive = false;
    int inst = -1;
    if (item < n_items) {
        inst = items ? items[item] : item;
        const RegionDesc rd = rt.rdesc[inst / n_init];
        int64_t g0 = rd.g0;
        MProblem pb;
        pb.na = rt.na + g0; pb.nb = rt.nb + g0; pb.invd = rt.invd + g0; pb.sq = rt.sq + g0;
        pb.L = rd.L;
        double Mr = (double)rd.M;
        int64_t NI = st.NI;
        pb.t0 = st.ES[inst] / Mr; pb.t1 = st.ES[NI + inst] / Mr; pb.t2 = st.ES[2 * NI + inst] / Mr;
        pb.mode = mode; pb.evals = 0;
        double x0[3] = {st.phi[inst], st.phi[NI + inst], st.phi[2 * NI + inst]}, sol[3];
        int nit; bool threw;
        bool conv = trust_region_solve(pb, x0, sol, &nit, &threw);
        if (threw || !allfinite3(sol)) st.numerr[inst] = 1;
        double nx[3];
        if (threw || !conv) { nx[0] = x0[0]; nx[1] = x0[1]; nx[2] = x0[2]; }
        else {
            nx[0] = fmax(-amax, fmin(sol[0], amax));
            nx[1] = fmax(-bmax, fmin(sol[1], bmax));
            nx[2] = fmax(-cmax, fmin(sol[2], cmax));
        }
        double d0 = x0[0] - nx[0], d1 = x0[1] - nx[1], d2 = x0[2] - nx[2];
        bool cv = sqrt(d0 * d0 + d1 * d1 + d2 * d2) <= 3.0e-2;          // conv_check, em_algorithm.jl:18-23
        int i = st.iter[inst];
        bool hit = i >= max_em_iters;
        i += 1;
        st.phi[inst] = nx[0]; st.phi[NI + inst] = nx[1]; st.phi[2 * NI + inst] = nx[2];
        st.iter[inst] = i;
        st.cnt_em[inst] += 1; st.cnt_sol[inst] += nit; st.cnt_eval[inst] += pb.evals;
        if (cv || hit) st.status[inst] = (cv && i > 2) ? 1 : 2;
        else survive = true;
    }
    // ordered compaction inside the block, one atomic per block
    __shared__ int warp_cnt[M_THREADS / 32];
    __shared__ int base;
    unsigned bal = __ballot_sync(0xffffffffu, survive);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[w] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int k = 0; k < M_THREADS / 32; ++k) { int c = warp_cnt[k]; warp_cnt[k] = tot; tot += c; }
        base = tot ? atomicAdd(n_next, tot) : 0;
    }
    __syncthreads();
    if (survive) next_items[base + warp_cnt[w] + __popc(bal & ((1u << lane) - 1))] = inst;
}

// Analytic-mode M-step with lane-level work stealing.  Every loop trip performs exactly one derivative sweep per busy
// lane (the expensive, uniform part); the trust-region bookkeeping between sweeps is a per-lane state machine
// (NLsolve trust_region_: initial evaluation, then one evaluation per iteration) written branch-light — both phases and
// all three dogleg cases run the same instruction stream with selects — so lanes at different solver phases do not
// serialise.  A lane whose solve ends takes the next EM instance from a global queue.  The solver state is cold during a
// sweep and lives in shared memory ([field][thread], conflict-free), leaving the registers to the 20-value recursion.
enum { C_X0 = 0, C_X = 3, C_XOLD = 6, C_R = 9, C_J = 12, C_D = 18, C_P = 21, C_DELTA = 24, C_T = 25, C_N = 28 };

// Solves the symmetric 3x3 system J p = r (J as aa,ab,ac,bb,bc,cc) by LU with partial pivoting; pivots are inverted once.
__device__ __forceinline__ bool lu_solve3_sym(const double* Js, const double* r, double* out) {
    double A[3][3] = {{Js[0], Js[1], Js[2]}, {Js[1], Js[3], Js[4]}, {Js[2], Js[4], Js[5]}};
    double b[3] = {r[0], r[1], r[2]};
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        // partial pivoting by conditional row swaps (selects)
#pragma unroll
        for (int i = k + 1; i < 3; ++i) {
            bool sw = fabs(A[i][k]) > fabs(A[k][k]);
#pragma unroll
            for (int j = 0; j < 3; ++j) { double u = A[k][j], v = A[i][j]; A[k][j] = sw ? v : u; A[i][j] = sw ? u : v; }
            double u = b[k], v = b[i]; b[k] = sw ? v : u; b[i] = sw ? u : v;
        }
        ok = ok && (A[k][k] != 0.0);
        double inv = 1.0 / A[k][k];
        A[k][k] = inv;
#pragma unroll
        for (int i = k + 1; i < 3; ++i) {
            double m = A[i][k] * inv;
#pragma unroll
            for (int j = k + 1; j < 3; ++j) A[i][j] = fma(-m, A[k][j], A[i][j]);
            b[i] = fma(-m, b[k], b[i]);
        }
    }
    out[2] = b[2] * A[2][2];
    out[1] = (b[1] - A[1][2] * out[2]) * A[1][1];
    out[0] = (b[0] - A[0][1] * out[1] - A[0][2] * out[2]) * A[0][0];
    return ok;
}
// NLsolve dogleg! with all three cases evaluated and selected (no divergent branches).
__device__ __forceinline__ void dogleg3_sel(double* p, const double* r, const double* d, const double* Js, double delta) {
    double p_i[3];
    bool ok = lu_solve3_sym(Js, r, p_i);
#pragma unroll
    for (int i = 0; i < 3; ++i) p_i[i] = -p_i[i];
    const bool case1 = ok && (wnorm3(d, p_i) <= delta);
    double g[3], Jg[3], p_c[3];
    g[0] = (Js[0] * r[0] + Js[1] * r[1] + Js[2] * r[2]) / (d[0] * d[0]);
    g[1] = (Js[1] * r[0] + Js[3] * r[1] + Js[4] * r[2]) / (d[1] * d[1]);
    g[2] = (Js[2] * r[0] + Js[4] * r[1] + Js[5] * r[2]) / (d[2] * d[2]);
    Jg[0] = Js[0] * g[0] + Js[1] * g[1] + Js[2] * g[2];
    Jg[1] = Js[1] * g[0] + Js[3] * g[1] + Js[4] * g[2];
    Jg[2] = Js[2] * g[0] + Js[4] * g[1] + Js[5] * g[2];
    const double wg2 = wdot3(d, g, d, g), wg = sqrt(wg2);
    const double coef = -wg2 / sumabs2_3(Jg);
#pragma unroll
    for (int i = 0; i < 3; ++i) p_c[i] = coef * g[i];
    const double wpc2 = wdot3(d, p_c, d, p_c);
    const bool case2 = !ok || (sqrt(wpc2) >= delta);
    const double s2 = -delta / wg;
    double p_diff[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) p_diff[i] = p_i[i] - p_c[i];
    const double bq = 2 * wdot3(d, p_c, d, p_diff);
    const double aq = wdot3(d, p_diff, d, p_diff);
    const double tau = (-bq + sqrt(bq * bq - 4 * aq * (wpc2 - delta * delta))) / (2 * aq);
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = case1 ? p_i[i] : (case2 ? g[i] * s2 : fma(tau, p_diff[i], p_c[i]));
}

// Loop trip of a lane:  [take work] → [B: consume the pending (∇logZ, Hessian): accept/reject, radius, convergence, next
// trial point] → [S: one derivative sweep at the trial point].  ∇logZ and its Hessian at the final iterate of one M-step
// are exactly what the next M-step of the same instance needs first (they do not depend on the E-step), so they are kept
// in `gh` and the first sweep of every later M-step is skipped (bit-identical to recomputing).  A lane whose instance ends
// in B immediately takes the next one (two rounds per trip), so no sweep slot is wasted.
__global__ void __launch_bounds__(M_THREADS, WS_BLOCKS) k_mstep_ws(int n_items, const int* __restrict__ items, int n_init, RegionTables rt,
                                                                    EmState st, int max_em_iters, double amax, double bmax, double cmax,
                                                                    int* __restrict__ queue,
                                                                    unsigned char* __restrict__ alive /*[n_items] by list position*/,
                                                                    const int* __restrict__ n_dev) {
    if (n_dev) n_items = __ldg(n_dev);
    __shared__ double cold[C_N][M_THREADS];
    __shared__ double2 ring[RING_SLOTS * 2 * M_THREADS];
#define CS(i) cold[(i)][threadIdx.x]
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1;
    const double ftol = 1e-3, eta = 1e-4;
    const int64_t NI = st.NI;
    const SweepSite* tbl = nullptr;
    int L = 0, inst = -1, pos = -1, it = 0, evals = 0;
    bool init = true, have = false, drained = false, pending = false;
    double g[3], H[6];
    for (;;) {
#pragma unroll 1
        for (int round = 0; round < WS_ROUNDS; ++round) {
            // ---- take work (warp-aggregated fetch from the global queue)
            bool need = !have && !drained;
            unsigned m = __ballot_sync(0xffffffffu, need);
            if (m) {
                int leader = __ffs(m) - 1, base = 0;
                if (lane == leader) base = atomicAdd(queue, __popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (need) {
                    int item = base + __popc(m & lt_mask);
                    if (item < n_items) {
                        inst = items ? items[item] : item;
                        pos = item;
                        const RegionDesc rd = rt.rdesc[inst / n_init];
                        tbl = rt.sq + rd.g0;
                        L = rd.L;
                        double Mr = (double)rd.M;
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            CS(C_T + i) = st.ES[i * NI + inst] / Mr;
                            double v = st.phi[i * NI + inst];
                            CS(C_X0 + i) = v; CS(C_X + i) = v;
                        }
                        evals = 0; it = 0; init = true; have = true;
                        pending = st.gh_valid[inst] != 0;
                        if (pending) {
#pragma unroll
                            for (int i = 0; i < 3; ++i) g[i] = st.gh[i * NI + inst];
#pragma unroll
                            for (int i = 0; i < 6; ++i) H[i] = st.gh[(3 + i) * NI + inst];
                        }
                    } else drained = true;
                }
            }
            // ---- B: trust-region bookkeeping (NLsolve trust_region_), both phases in one instruction stream
            if (have && pending) {
                pending = false;
                double x[3] = {CS(C_X), CS(C_X + 1), CS(C_X + 2)};
                double fval[3] = {g[0] - CS(C_T), g[1] - CS(C_T + 1), g[2] - CS(C_T + 2)};
                double r[3] = {CS(C_R), CS(C_R + 1), CS(C_R + 2)};
                double p[3] = {CS(C_P), CS(C_P + 1), CS(C_P + 2)};
                double d[3] = {CS(C_D), CS(C_D + 1), CS(C_D + 2)};
                double Js[6] = {CS(C_J), CS(C_J + 1), CS(C_J + 2), CS(C_J + 3), CS(C_J + 4), CS(C_J + 5)};
                double delta = CS(C_DELTA);
                double rp[3] = {Js[0] * p[0] + Js[1] * p[1] + Js[2] * p[2] + r[0], Js[1] * p[0] + Js[3] * p[1] + Js[4] * p[2] + r[1],
                                Js[2] * p[0] + Js[4] * p[1] + Js[5] * p[2] + r[2]};
                double rho = (sumabs2_3(r) - sumabs2_3(fval)) / (sumabs2_3(r) - sumabs2_3(rp));
                const bool accept = init || (rho > eta);
                bool x_conv = false;
                if (accept) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) r[i] = fval[i];
#pragma unroll
                    for (int i = 0; i < 6; ++i) Js[i] = H[i];
                    double n0 = sqrt(H[0] * H[0] + H[1] * H[1] + H[2] * H[2]);
                    double n1 = sqrt(H[1] * H[1] + H[3] * H[3] + H[4] * H[4]);
                    double n2 = sqrt(H[2] * H[2] + H[4] * H[4] + H[5] * H[5]);
                    d[0] = init ? (n0 == 0.0 ? 1.0 : n0) : fmax(0.1 * d[0], n0);
                    d[1] = init ? (n1 == 0.0 ? 1.0 : n1) : fmax(0.1 * d[1], n1);
                    d[2] = init ? (n2 == 0.0 ? 1.0 : n2) : fmax(0.1 * d[2], n2);
                    double dx[3] = {x[0] - CS(C_XOLD), x[1] - CS(C_XOLD + 1), x[2] - CS(C_XOLD + 2)};
                    x_conv = !init && (norminf3(dx) <= 0.0);
                    if (!init || evals > 0) {          // remember the evaluation at the accepted point (unless it came from gh)
#pragma unroll
                        for (int i = 0; i < 3; ++i) st.gh[i * NI + inst] = g[i];
#pragma unroll
                        for (int i = 0; i < 6; ++i) st.gh[(3 + i) * NI + inst] = H[i];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 3; ++i) x[i] -= p[i];
                }
                if (init) {
                    delta = wnorm3(d, x);
                    if (delta == 0.0) delta = 1.0;
                } else {
                    const double wdp = wnorm3(d, p);                 // with the scaling vector already updated (as NLsolve does)
                    if (rho < 0.1) delta = delta / 2;
                    else if (rho >= 0.9) delta = 2 * wdp;
                    else if (rho >= 0.5) delta = fmax(delta, 2 * wdp);
                }
                const bool threw = init && !allfinite3(fval);
                const bool conv = accept && !threw && (x_conv || norminf3(r) <= ftol);
                const bool stopped = !allfinite3(x) || anynan3(fval);      // NLsolve stops on NaN; a non-finite x gives NaN at its next evaluation
                int code = 0;
                if (threw) code = 2;
                else if (conv) code = 1;
                else if (stopped || (!init && it >= 20)) code = 2;
                if (code == 0) {
                    it = init ? 1 : it + 1;
                    init = false;
                    dogleg3_sel(p, r, d, Js, delta);
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        CS(C_XOLD + i) = x[i]; CS(C_X + i) = x[i] + p[i]; CS(C_P + i) = p[i]; CS(C_R + i) = r[i]; CS(C_D + i) = d[i];
                    }
#pragma unroll
                    for (int i = 0; i < 6; ++i) CS(C_J + i) = Js[i];
                    CS(C_DELTA) = delta;
                } else {
                    // em_iter tail (:172-181) + em_inst bookkeeping (:327-341)
                    double x0[3] = {CS(C_X0), CS(C_X0 + 1), CS(C_X0 + 2)}, nx[3];
                    bool gh_ok;
                    if (threw || stopped) st.numerr[inst] = 1;
                    if (code == 2) {
                        nx[0] = x0[0]; nx[1] = x0[1]; nx[2] = x0[2];
                        gh_ok = init && !threw;               // ϕ stays at x0: gh still describes it only if no step was accepted
                    } else {
                        nx[0] = fmax(-amax, fmin(x[0], amax));
                        nx[1] = fmax(-bmax, fmin(x[1], bmax));
                        nx[2] = fmax(-cmax, fmin(x[2], cmax));
                        gh_ok = (nx[0] == x[0]) && (nx[1] == x[1]) && (nx[2] == x[2]);   // gh was evaluated at x, not at the clamped point
                    }
                    double d0 = x0[0] - nx[0], d1 = x0[1] - nx[1], d2 = x0[2] - nx[2];
                    bool cv = sqrt(d0 * d0 + d1 * d1 + d2 * d2) <= 3.0e-2;
                    int i = st.iter[inst];
                    bool hit = i >= max_em_iters;
                    i += 1;
                    st.phi[inst] = nx[0]; st.phi[NI + inst] = nx[1]; st.phi[2 * NI + inst] = nx[2];
                    st.iter[inst] = i;
                    st.cnt_em[inst] += 1; st.cnt_sol[inst] += init ? 0 : it; st.cnt_eval[inst] += evals;
                    st.gh_valid[inst] = gh_ok ? 1 : 0;
                    bool done = cv || hit;
                    if (done) st.status[inst] = (cv && i > 2) ? 1 : 2;
                    alive[pos] = done ? 0 : 1;
                    have = false;
                }
            }
            if (!__any_sync(0xffffffffu, !have && !drained)) break;
        }
        if (!__any_sync(0xffffffffu, have)) break;
        // ---- S: one derivative sweep at the current (trial) point of every busy lane
        if (have) {
            eval_grad_hess<true>(tbl, L, CS(C_X), CS(C_X + 1), CS(C_X + 2), g, H, ring);
            evals += 1;
            pending = true;
        }
    }
#undef CS
}

// Order-preserving compaction of the active list (keeps instances sorted by chain length so that the lanes of an
// M-step warp walk chains of similar length): per-block counts → scan of the counts → scatter.
constexpr int CP_THREADS = 256, CP_ITEMS = 4, CP_TILE = CP_THREADS * CP_ITEMS;

__global__ void __launch_bounds__(CP_THREADS) k_compact_count(int n, const unsigned char* __restrict__ alive, int* __restrict__ blk_cnt,
                                                               const int* __restrict__ n_dev) {
    __shared__ int wsum[CP_THREADS / 32];
    if (n_dev) n = __ldg(n_dev);
    int base = blockIdx.x * CP_TILE + threadIdx.x * CP_ITEMS;
    int c = 0;
#pragma unroll
    for (int k = 0; k < CP_ITEMS; ++k) c += (base + k < n) ? alive[base + k] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < CP_THREADS / 32; ++k) t += wsum[k];
        blk_cnt[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(1024) k_compact_scan(int nblk, int* __restrict__ blk_cnt /*in: counts, out: exclusive offsets*/,
                                                        int* __restrict__ total) {
    __shared__ int sm[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nblk; b0 += 1024) {
        int i = b0 + threadIdx.x;
        int v = (i < nblk) ? blk_cnt[i] : 0;
        sm[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int add = (threadIdx.x >= o) ? sm[threadIdx.x - o] : 0;
            __syncthreads();
            sm[threadIdx.x] += add;
            __syncthreads();
        }
        if (i < nblk) blk_cnt[i] = carry + sm[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += sm[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(CP_THREADS) k_compact_scatter(int n, const unsigned char* __restrict__ alive, const int* __restrict__ items,
                                                                 const int* __restrict__ blk_off, int* __restrict__ out,
                                                                 const int* __restrict__ n_dev) {
    __shared__ int wsum[CP_THREADS / 32];
    if (n_dev) n = __ldg(n_dev);
    int base = blockIdx.x * CP_TILE + threadIdx.x * CP_ITEMS;
    int f[CP_ITEMS], c = 0;
#pragma unroll
    for (int k = 0; k < CP_ITEMS; ++k) { f[k] = (base + k < n) ? alive[base + k] : 0; c += f[k]; }
    int incl = c;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    int woff = 0;
    for (int k = 0; k < w; ++k) woff += wsum[k];
    int dst = blk_off[blockIdx.x] + woff + incl - c;
#pragma unroll
    for (int k = 0; k < CP_ITEMS; ++k) if (f[k]) out[dst++] = items[base + k];
}

// Single solve for the test hook cpn_mstep_batch.
__global__ void k_mstep_once(int64_t R, RegionTables rt, const int64_t* __restrict__ n_reads, const double* __restrict__ ES,
                             const double* __restrict__ phi_init, int mode, double* __restrict__ sol_out, int* __restrict__ conv_out,
                             int64_t* __restrict__ counters) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= R) return;
    int64_t g0 = rt.grp_off[r];
    MProblem pb;
    pb.na = rt.na + g0; pb.nb = rt.nb + g0; pb.invd = rt.invd + g0; pb.sq = rt.sq + g0;
    pb.L = (int)(rt.grp_off[r + 1] - g0);
    double Mr = (double)n_reads[r];
    pb.t0 = ES[3 * r] / Mr; pb.t1 = ES[3 * r + 1] / Mr; pb.t2 = ES[3 * r + 2] / Mr;
    pb.mode = mode; pb.evals = 0;
    double x0[3] = {phi_init[3 * r], phi_init[3 * r + 1], phi_init[3 * r + 2]}, sol[3];
    int nit; bool threw;
    bool conv = trust_region_solve(pb, x0, sol, &nit, &threw);
    sol_out[3 * r] = sol[0]; sol_out[3 * r + 1] = sol[1]; sol_out[3 * r + 2] = sol[2];
    conv_out[r] = (conv && !threw) ? 1 : 0;
    if (counters) { counters[2 * r] = nit; counters[2 * r + 1] = pb.evals; }
}

// EM starts: phi0 [NI][3] from the caller (the host RNG owns gen_ϕ0s), or — phi0 == null — regenerated from the seed with the
// counter-based stream of cpn_philox.h: unit u is region unit_base + u of the call's batch (plain fit), or the (region,
// permutation, half) its key names (two-sample refits).
__global__ void k_init_state(int64_t NI, int n_init, const double* __restrict__ phi0 /*[NI][3] or null*/, uint64_t seed, int64_t unit_base,
                             const int2* __restrict__ ukey /*[U] (region, 2·perm + half) or null*/, EmState st, int* __restrict__ items,
                             const int* __restrict__ order /*[R] region order or null*/) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= NI) return;
    if (phi0) {
        st.phi[i] = phi0[3 * i]; st.phi[st.NI + i] = phi0[3 * i + 1]; st.phi[2 * st.NI + i] = phi0[3 * i + 2];
    } else {
        const int64_t u = i / n_init;
        const uint32_t sidx = (uint32_t)(i - u * n_init);
        double p[3];
        if (ukey) { const int2 k = ukey[u]; cpn_gen_phi0(seed, (uint32_t)k.x, (uint32_t)k.y >> 1, (uint32_t)k.y & 1u, sidx, p); }
        else cpn_gen_phi0(seed, (uint32_t)(unit_base + u), CPN_RNG_NO_PERM, 0u, sidx, p);
        st.phi[i] = p[0]; st.phi[st.NI + i] = p[1]; st.phi[2 * st.NI + i] = p[2];
    }
    st.iter[i] = 1; st.status[i] = 0; st.cnt_em[i] = 0; st.cnt_sol[i] = 0; st.cnt_eval[i] = 0;
    st.gh_valid[i] = 0; st.numerr[i] = 0;
    if (order) {
        int64_t slot = i / n_init, s = i % n_init;
        items[i] = (int)((int64_t)order[slot] * n_init + s);
    } else items[i] = (int)i;
}

// pick_ϕ (em_algorithm.jl:35-53): first start with the strictly largest Q; ϕ̂ = [] (NaN here) if all −Inf (:385)
__global__ void k_pick(int64_t R, int n_init, EmState st, const double* __restrict__ Qs, double* __restrict__ phi_hat,
                       double* __restrict__ Q, int* __restrict__ status, int64_t* __restrict__ counters,
                       double* __restrict__ phi_starts, double* __restrict__ Q_starts) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= R) return;
    double best = -cpn_inf();
    int bi = -1;
    int64_t ce = 0, cs = 0, cv = 0;
    bool numerr = false;
    for (int s = 0; s < n_init; ++s) {
        int64_t i = r * n_init + s;
        double q = Qs[i];
        if (q > best) { best = q; bi = s; }
        numerr = numerr || st.numerr[i] != 0;
        ce += st.cnt_em[i]; cs += st.cnt_sol[i]; cv += st.cnt_eval[i];
        if (phi_starts) {
            phi_starts[3 * i] = st.phi[i]; phi_starts[3 * i + 1] = st.phi[st.NI + i]; phi_starts[3 * i + 2] = st.phi[2 * st.NI + i];
        }
        if (Q_starts) Q_starts[i] = q;
    }
    if (bi >= 0) {
        int64_t i = r * n_init + bi;
        phi_hat[3 * r] = st.phi[i]; phi_hat[3 * r + 1] = st.phi[st.NI + i]; phi_hat[3 * r + 2] = st.phi[2 * st.NI + i];
    } else {
        phi_hat[3 * r] = phi_hat[3 * r + 1] = phi_hat[3 * r + 2] = cpn_nan();
    }
    Q[r] = best;
    if (status) status[r] = bi >= 0 ? bi : (numerr ? CPN_FIT_NUMERICAL : CPN_FIT_NONE);
    if (counters) { counters[4 * r] = ce; counters[4 * r + 1] = cs; counters[4 * r + 2] = cv; counters[4 * r + 3] = ce; }
}

// Algorithmic work done by the EM loop (roofline accounting): Σ_i passes_i·M·L, Σ_i sweeps_i·L, Σ_converged M·L, instance count.
__global__ void k_work_sum(int64_t NI, int n_init, const RegionDesc* __restrict__ rdesc, EmState st, double* __restrict__ out /*[4]*/) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double w0 = 0, w1 = 0, w2 = 0, w3 = 0;
    if (i < NI) {
        const RegionDesc rd = rdesc[i / n_init];
        const double ml = (double)rd.M * (double)rd.L;
        w0 = st.cnt_em[i] * ml; w1 = (double)st.cnt_eval[i] * rd.L; w2 = st.status[i] == 1 ? ml : 0.0; w3 = 1.0;
    }
    w0 = cpn_warp_sum(w0); w1 = cpn_warp_sum(w1); w2 = cpn_warp_sum(w2); w3 = cpn_warp_sum(w3);
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, w0); atomicAdd(out + 1, w1); atomicAdd(out + 2, w2); atomicAdd(out + 3, w3); }
}

__global__ void k_pack_phi(int64_t R, const double* __restrict__ phi /*[R][3]*/, double* __restrict__ out /*[3][R]*/) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= R) return;
    out[r] = phi[3 * r]; out[R + r] = phi[3 * r + 1]; out[2 * R + r] = phi[3 * r + 2];
}
__global__ void k_unpack3(int64_t R, const double* __restrict__ in /*[3][R]*/, double* __restrict__ out /*[R][3]*/) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= R) return;
    out[3 * r] = in[r]; out[3 * r + 1] = in[R + r]; out[3 * r + 2] = in[2 * R + r];
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct PackedRegions {
    DevIn<int64_t> grp_off, read_off;
    DevIn<double> Nl, rho, dl, lu, lm;
    DevBuf<int64_t> obs_off, ew_off, rb_off;
    DevBuf<double> na, nb, invd, ew, ewi, rb;
    DevBuf<SweepSite> sq;
    DevBuf<RegionDesc> rdesc;
    std::vector<int64_t> h_grp_off, h_read_off;
    int64_t R = 0, sumL = 0, n_obs = 0;
    RegionTables rt{};
};

inline unsigned blocks_for(int64_t n, int per_block) { return (unsigned)((n + per_block - 1) / per_block); }

static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
// Once per device: k_estep_pair asks for more than the 48 KB static shared-memory limit, and EP_BLOCKS of its blocks only fit an SM
// when the L1/shared split is at its shared-memory end (left to itself the driver may pick a smaller carveout and keep fewer
// blocks resident).  CPN_EP_CARVEOUT: percent of the maximum, -1 = the driver's choice.
static cudaError_t ep_prepare(cpn_ctx* ctx) {
    if (ctx->ep_attr_set) return cudaSuccess;
    cudaError_t e;
    const int carve = env_int("CPN_EP_CARVEOUT", 100);
    if ((e = cudaFuncSetAttribute(k_estep_pair<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EP_SMEM)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_estep_pair<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EP_SMEM)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_estep_pair_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EPW_SMEM)) != cudaSuccess) return e;
    if (carve >= 0) {
        if ((e = cudaFuncSetAttribute(k_estep_pair<false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(k_estep_pair<true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(k_estep_pair_ws, cudaFuncAttributePreferredSharedMemoryCarveout, carve)) != cudaSuccess) return e;
    }
    {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_estep_pair_ws, E_WARPS * 32, EPW_SMEM) != cudaSuccess || nb < 1) nb = 1;
        ctx->epw_blocks = nb;
    }
    if (env_int("CPN_TRACE", 0)) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_estep_pair<false>, E_WARPS * 32, EP_SMEM);
        fprintf(stderr, "[cpn] k_estep_pair: %zu B shared memory per block, %d blocks per SM possible, carveout %d\n", (size_t)EP_SMEM, nb, carve);
    }
    ctx->ep_attr_set = true;
    return cudaSuccess;
}

// CPN_TRACE=1 prints host-side phase timestamps (ms since the first traced event) to stderr.
static void trace(const cpn_ctx* ctx, const char* what) {
    static const bool on = getenv("CPN_TRACE") != nullptr;
    if (!on) return;
    static const auto epoch = std::chrono::steady_clock::now();
    double t = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - epoch).count();
    uint64_t reserved = 0, used = 0;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
    }
    fprintf(stderr, "[cpn trace] %p %-10s %9.2f  pool reserved %7.1f MB used %7.1f MB\n", (const void*)ctx, what, t, reserved / 1048576.0,
            used / 1048576.0);
}

// Uploads (or borrows) the CSR inputs and builds the per-group and per-read device tables.
int pack_regions(cpn_ctx* ctx, PackedRegions& pk, bool on_device, bool with_reads, int64_t R, const int64_t* grp_off,
                 const double* Nl, const double* rho_l, const double* d_l, const int64_t* read_off, const double* lu, const double* lm,
                 int offs_mode = -1 /*-1: offsets live where the data lives; 0: offsets on the host, data per on_device*/,
                 bool packed = false /*lu = log-likelihood ratios, lm = per-read Σ_obs log p(y|−1) (cpn_pack_emissions)*/) {
    cudaStream_t s = ctx->stream;
    const bool offs_dev = offs_mode < 0 ? on_device : offs_mode != 0;
    pk.R = R;
    pk.h_grp_off.resize(R + 1);
    pk.h_read_off.assign(R + 1, 0);
    if (offs_dev) {
        CPN_CUDA(ctx, cudaMemcpyAsync(pk.h_grp_off.data(), grp_off, (R + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        if (with_reads) CPN_CUDA(ctx, cudaMemcpyAsync(pk.h_read_off.data(), read_off, (R + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        CPN_CUDA(ctx, cpn_wait_stream(ctx, s));
    } else {
        std::memcpy(pk.h_grp_off.data(), grp_off, (R + 1) * sizeof(int64_t));
        if (with_reads) std::memcpy(pk.h_read_off.data(), read_off, (R + 1) * sizeof(int64_t));
    }
    CPN_ARG(ctx, pk.h_grp_off[0] == 0, "grp_off[0] must be 0");
    std::vector<int64_t> h_obs(R + 1), h_ew(R + 1), h_rb(R + 1);
    h_obs[0] = h_ew[0] = h_rb[0] = 0;
    for (int64_t r = 0; r < R; ++r) {
        int64_t L = pk.h_grp_off[r + 1] - pk.h_grp_off[r];
        int64_t M = pk.h_read_off[r + 1] - pk.h_read_off[r];
        CPN_ARG(ctx, L >= 2, "every region needs L >= 2 CpG groups (nanopore.jl:326)");
        CPN_ARG(ctx, !with_reads || M >= 1, "every region needs at least one read");
        int64_t nch = (M + 31) / 32;
        h_obs[r + 1] = h_obs[r] + M * L;
        h_ew[r + 1] = h_ew[r] + nch * L * EW_ROW;
        h_rb[r + 1] = h_rb[r] + nch * 32;
    }
    pk.sumL = pk.h_grp_off[R];
    pk.n_obs = h_obs[R];
    trace(ctx, "pk-begin");
    CPN_CUDA(ctx, pk.grp_off.bind(grp_off, R + 1, offs_dev, s));
    CPN_CUDA(ctx, pk.Nl.bind(Nl, pk.sumL, on_device, s));
    CPN_CUDA(ctx, pk.rho.bind(rho_l, pk.sumL, on_device, s));
    CPN_CUDA(ctx, pk.dl.bind(d_l, pk.sumL - R, on_device, s));
    CPN_CUDA(ctx, pk.na.alloc(pk.sumL, s));
    CPN_CUDA(ctx, pk.nb.alloc(pk.sumL, s));
    CPN_CUDA(ctx, pk.invd.alloc(pk.sumL, s));
    CPN_CUDA(ctx, pk.sq.alloc(pk.sumL, s));
    { KTimer kt(ctx, CPN_K_PREP);
      k_prep_sites<<<blocks_for(R * 32, 256), 256, 0, s>>>(R, pk.grp_off.p, pk.Nl.p, pk.rho.p, pk.dl.p, pk.na.p, pk.nb.p, pk.invd.p, pk.sq.p); }
    pk.rt.grp_off = pk.grp_off.p; pk.rt.na = pk.na.p; pk.rt.nb = pk.nb.p; pk.rt.invd = pk.invd.p; pk.rt.sq = pk.sq.p;
    if (with_reads) {
        trace(ctx, "pk-sites");
        // tables first, raw emissions last: they sit on top of the workspace stack and are popped once packed
        std::vector<RegionDesc> h_desc(R);
        for (int64_t r = 0; r < R; ++r) {
            h_desc[r].g0 = pk.h_grp_off[r]; h_desc[r].ew_off = h_ew[r]; h_desc[r].rb_off = h_rb[r];
            h_desc[r].L = (int32_t)(pk.h_grp_off[r + 1] - pk.h_grp_off[r]); h_desc[r].M = (int32_t)(pk.h_read_off[r + 1] - pk.h_read_off[r]);
            h_desc[r].ridx = -1; h_desc[r].pad = 0; h_desc[r].llrmax = 0.f; h_desc[r].idmax = 0.f;
        }
        CPN_CUDA(ctx, pk.read_off.bind(read_off, R + 1, offs_dev, s));
        CPN_CUDA(ctx, pk.obs_off.upload(h_obs.data(), R + 1, s));
        CPN_CUDA(ctx, pk.ew_off.upload(h_ew.data(), R + 1, s));
        CPN_CUDA(ctx, pk.rb_off.upload(h_rb.data(), R + 1, s));
        CPN_CUDA(ctx, pk.rdesc.upload(h_desc.data(), R, s));
        pk.rt.rdesc = pk.rdesc.p;
        CPN_CUDA(ctx, pk.ew.alloc(h_ew[R], s));
        CPN_CUDA(ctx, pk.ewi.alloc(h_ew[R], s));
        CPN_CUDA(ctx, pk.rb.alloc(h_rb[R], s));
        CPN_CUDA(ctx, pk.lu.bind(lu, pk.n_obs, on_device, s));
        trace(ctx, "pk-lu");
        CPN_CUDA(ctx, pk.lm.bind(lm, packed ? pk.h_read_off[R] : pk.n_obs, on_device, s));
        trace(ctx, "pk-lm");
        { KTimer kt(ctx, CPN_K_PREP);
          k_prep_emis<<<blocks_for(R * 32, 128), 128, 0, s>>>(R, pk.grp_off.p, pk.read_off.p, pk.obs_off.p, pk.lu.p, pk.lm.p, packed, pk.ew_off.p,
                                                              pk.rb_off.p, pk.invd.p, pk.ew.p, pk.ewi.p, pk.rb.p, pk.rdesc.p); }
        // the raw emissions are not needed once packed: hand the memory back (stream-ordered) before the EM loop
        pk.lm.own.release();
        pk.lu.own.release();
        trace(ctx, "pk-enq");
        CPN_CUDA(ctx, cpn_wait_stream(ctx, s));   // h_* staging vectors go out of scope
        trace(ctx, "pk-sync");
        pk.rt.read_off = pk.read_off.p; pk.rt.ew_off = pk.ew_off.p; pk.rt.rb_off = pk.rb_off.p; pk.rt.ew = pk.ew.p; pk.rt.ewi = pk.ewi.p; pk.rt.rb = pk.rb.p;
    }
    CPN_CUDA(ctx, cudaGetLastError());
    return CPN_OK;
}

// Optional second stage of cpn_analyze_batch: summaries of the fitted per-CpG chains.
struct SummaryArgs {
    const int64_t *cpg_off, *sub_off, *sub_p, *sub_q;
    const double *rho_n, *d_n;
    double *logZ, *ex, *exx, *log_g, *e_log_g1, *e_log_g2, *mml, *nme;
};

// Pinned count mailbox and one event per EM iteration, owned by the context and grown on demand.
static int ctx_iter_resources(cpn_ctx* ctx, int n) {
    if (ctx->h_counts_cap < n) {
        if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
        ctx->h_counts = nullptr; ctx->h_counts_cap = 0;
        CPN_CUDA(ctx, cudaHostAlloc((void**)&ctx->h_counts, (size_t)n * sizeof(int), cudaHostAllocDefault));
        ctx->h_counts_cap = n;
    }
    while ((int)ctx->it_ev.size() < n) {
        cudaEvent_t e;
        CPN_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming | (cpn_spin_wait() ? 0 : cudaEventBlockingSync)));
        ctx->it_ev.push_back(e);
    }
    return CPN_OK;
}

// Inputs of one batch (or one chunk of a batch) resident on the device: the packed region tables and the EM starts.
struct StagedBatch {
    CpnArena* arena = nullptr;   // workspace the tables live in (reset by whoever hands it out)
    PackedRegions pk;
    DevIn<double> phi0;
    // optional views (two-sample permutation refits): the EM runs over V views of the staged regions instead of the regions
    int64_t V = 0;
    DevBuf<RegionDesc> vdesc;    // [V]
    DevBuf<int32_t> ridx;        // read-id rows
    std::vector<int32_t> h_vL;   // chain length per view (ordering)
    // EM starts regenerated on the device (phi0 == null): seed and, for plain fits, the position of unit 0 in the call's
    // batch; views carry their (region, 2·perm + half) key
    bool seeded = false;
    uint64_t seed = 0;
    int64_t unit_base = 0;
    DevBuf<int2> vkey;           // [V]
};

int fit_stage(cpn_ctx* ctx, StagedBatch& sb, bool on_device, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l,
              const double* d_l, const int64_t* read_off, const double* lu, const double* lm, const double* phi0, int32_t n_init,
              int offs_mode = -1, int64_t n_units = -1 /*EM units the starts are for; default: the R regions*/, bool packed = false) {
    CPN_CUDA(ctx, cudaSetDevice(ctx->device));
    CpnArenaScope scope(sb.arena);
    // ϕ0 first: the raw emissions must stay on top of the workspace stack
    sb.seeded = phi0 == nullptr;
    if (!sb.seeded) CPN_CUDA(ctx, sb.phi0.bind(phi0, (n_units < 0 ? R : n_units) * n_init * 3, on_device, ctx->stream));
    return pack_regions(ctx, sb.pk, on_device, true, R, grp_off, Nl, rho_l, d_l, read_off, lu, lm, offs_mode, packed);
}

// `pre` = inputs staged earlier (possibly on another context's stream, already synchronised); the input pointers are then unused.
int fit_impl(cpn_ctx* ctx, bool on_device, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
             const int64_t* read_off, const double* lu, const double* lm, const double* phi0, int32_t n_init, int32_t max_em_iters,
             double amax, double bmax, double cmax, int32_t mstep_mode, double* phi_hat, double* Q, int32_t* status, int64_t* counters,
             double* phi_starts, double* Q_starts, const SummaryArgs* sa = nullptr, StagedBatch* pre = nullptr, int offs_mode = -1,
             bool packed = false) {
    const bool offs_dev = offs_mode < 0 ? on_device : offs_mode != 0;
    CPN_ARG(ctx, R >= 0 && n_init >= 1 && max_em_iters >= 1, "R >= 0, n_init >= 1, max_em_iters >= 1");
    CPN_ARG(ctx, pre || (grp_off && Nl && rho_l && d_l && read_off && lu && lm), "null pointer");
    CPN_ARG(ctx, phi_hat && Q, "null pointer");
    CPN_ARG(ctx, mstep_mode == CPN_MSTEP_FD || mstep_mode == CPN_MSTEP_ANALYTIC, "mstep_mode");
    CPN_ARG(ctx, R * (int64_t)n_init < (int64_t)2147483647, "R*n_init must fit in int32");
    CPN_CUDA(ctx, cudaSetDevice(ctx->device));
    LaunchCounter lc(ctx);
    if (R == 0) { ctx->last_ms = 0; return CPN_OK; }
    cudaStream_t s = ctx->stream;
    if (!pre) CPN_CUDA(ctx, cudaEventRecord(ctx->ev0, s));      // with `pre` the caller owns the timing events
    {
        StagedBatch local;
        StagedBatch* sb = pre ? pre : &local;
        trace(ctx, "begin");
        CPN_CUDA(ctx, ctx->arena.reset());
        if (!pre) {
            CPN_CUDA(ctx, ctx->stage_arena.reset());
            local.arena = &ctx->stage_arena;
            local.seed = ctx->seed;
            local.unit_base = ctx->region_base;
            int rc = fit_stage(ctx, local, on_device, R, grp_off, Nl, rho_l, d_l, read_off, lu, lm, phi0, n_init, -1, -1, packed);
            if (rc) return rc;
        }
        PackedRegions& pk = sb->pk;
        DevIn<double>& d_phi0 = sb->phi0;
        CpnArenaScope scope(&ctx->arena);
        // EM units: the staged regions, or views of them
        const bool views = sb->V > 0;
        CPN_ARG(ctx, !(views && sa), "summaries of views are computed by the caller");
        const int64_t U = views ? sb->V : R;
        RegionTables rt = pk.rt;
        if (views) { rt.rdesc = sb->vdesc.p; rt.ridx = sb->ridx.p; }
        const int64_t NI = U * n_init;
        DevBuf<double> phi, ES, Qs;
        DevBuf<int> iter, stat, c_em, c_sol, c_ev, items_a, items_b, order;
        CPN_CUDA(ctx, phi.alloc(3 * NI, s)); CPN_CUDA(ctx, ES.alloc(3 * NI, s)); CPN_CUDA(ctx, Qs.alloc(NI, s));
        CPN_CUDA(ctx, iter.alloc(NI, s)); CPN_CUDA(ctx, stat.alloc(NI, s)); CPN_CUDA(ctx, c_em.alloc(NI, s));
        CPN_CUDA(ctx, c_sol.alloc(NI, s)); CPN_CUDA(ctx, c_ev.alloc(NI, s));
        CPN_CUDA(ctx, items_a.alloc(NI, s)); CPN_CUDA(ctx, items_b.alloc(NI, s));
        DevBuf<unsigned char> alive;
        DevBuf<int> blk_cnt;
        DevBuf<ItemDesc> idesc;
        DevBuf<PairDesc> pdesc;
        DevBuf<double> phil;
        DevBuf<int> pair_cnt;
        // two instances per E-step warp (k_estep_pair) whenever the active list keeps a region's starts together, i.e. with the
        // order-preserving compaction of the analytic M-step; CPN_ESTEP_PAIRS=0 selects the one-instance kernel
        const bool pairs = mstep_mode == CPN_MSTEP_ANALYTIC && env_int("CPN_ESTEP_PAIRS", 1) != 0;
        const bool pair_ws = env_int("CPN_ESTEP_WS", 1) != 0;      // resident grid + work queue + prefetched head (plain regions)
        if (pairs) {
            CPN_CUDA(ctx, ep_prepare(ctx));
            CPN_CUDA(ctx, pdesc.alloc(NI, s));
            CPN_CUDA(ctx, phil.alloc(6 * NI, s));
            CPN_CUDA(ctx, pair_cnt.alloc(4 * (size_t)max_em_iters, s));          // per iteration: pairs with c ≥ 0, pairs with c < 0, queue head of k_estep_pair_ws
            CPN_CUDA(ctx, cudaMemsetAsync(pair_cnt.p, 0, 4 * (size_t)max_em_iters * sizeof(int), s));
        } else {
            CPN_CUDA(ctx, idesc.alloc(NI, s));
            CPN_CUDA(ctx, phil.alloc(3 * NI, s));
        }
        CPN_CUDA(ctx, alive.alloc(NI, s));
        CPN_CUDA(ctx, blk_cnt.alloc((NI + CP_TILE - 1) / CP_TILE, s));
        int ws_blocks_per_sm = 1;
                CPN_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ws_blocks_per_sm, k_mstep_ws, M_THREADS, 0));
        if (ws_blocks_per_sm < 1) ws_blocks_per_sm = 1;
        // Regions ordered by group count (counting sort) so the 32 lanes of an M-step warp walk chains of
        // similar length.
        std::vector<int> h_order(U);
        {
            auto len = [&](int64_t u) -> int64_t { return views ? sb->h_vL[u] : pk.h_grp_off[u + 1] - pk.h_grp_off[u]; };
            int64_t Lmax = 0;
            for (int64_t u = 0; u < U; ++u) Lmax = std::max(Lmax, len(u));
            std::vector<int64_t> cnt(Lmax + 2, 0);
            for (int64_t u = 0; u < U; ++u) cnt[len(u) + 1]++;
            for (int64_t l = 0; l <= Lmax; ++l) cnt[l + 1] += cnt[l];
            const bool desc = env_int("CPN_ORDER_DESC", 1) != 0;
            for (int64_t u = 0; u < U; ++u) {
                int64_t pos = cnt[len(u)]++;
                h_order[desc ? U - 1 - pos : pos] = (int)u;     // longest chains first: the kernels' tails are made of short ones
            }
        }
        CPN_CUDA(ctx, order.upload(h_order.data(), U, s));
        DevBuf<double> gh;
        DevBuf<unsigned char> gh_valid, numerr;
        CPN_CUDA(ctx, gh.alloc(9 * NI, s));
        CPN_CUDA(ctx, gh_valid.alloc(NI, s));
        CPN_CUDA(ctx, numerr.alloc(NI, s));
        EmState st{phi.p, ES.p, iter.p, stat.p, c_em.p, c_sol.p, c_ev.p, NI, gh.p, gh_valid.p, numerr.p};
        { KTimer kt(ctx, CPN_K_PREP);
          k_init_state<<<blocks_for(NI, 256), 256, 0, s>>>(NI, n_init, sb->seeded ? nullptr : d_phi0.p, sb->seed, sb->unit_base,
                                                           views ? sb->vkey.p : nullptr, st, items_a.p, order.p); }
        trace(ctx, "em-start");
        // The host never waits for the device inside the EM loop.  The number of instances still active after iteration i
        // lives in counts[i+1] on the device (kernels of iteration i+1 read it there); the host sizes its grids from the
        // newest count that has already arrived in pinned memory — an upper bound, since the active list only shrinks —
        // and runs at most EM_LAG iterations ahead.  Launch and read-back latencies (which grow several-fold while another
        // stream saturates PCIe with uploads) stay off the critical path.
        constexpr int EM_LAG = 2;
        DevBuf<int> counts;                                  // [max_em_iters + 1][2]: {active count, M-step queue head}
        CPN_CUDA(ctx, counts.alloc(2 * ((size_t)max_em_iters + 1), s));
        CPN_CUDA(ctx, cudaMemsetAsync(counts.p, 0, 2 * ((size_t)max_em_iters + 1) * sizeof(int), s));
        int rc_h = ctx_iter_resources(ctx, max_em_iters + 1);
        if (rc_h) return rc_h;
        volatile int* h_counts = ctx->h_counts;
        h_counts[0] = (int)NI;
        CPN_CUDA(ctx, cudaMemcpyAsync(counts.p, (const void*)ctx->h_counts, sizeof(int), cudaMemcpyHostToDevice, s));
        int known = 0;
        int* cur = items_a.p;
        int* nxt = items_b.p;
        for (int it = 0; it < max_em_iters; ++it) {
            if (it - EM_LAG > known) { CPN_CUDA(ctx, cudaEventSynchronize(ctx->it_ev[it - EM_LAG])); known = it - EM_LAG; }
            while (known < it && cudaEventQuery(ctx->it_ev[known + 1]) == cudaSuccess) ++known;
            const int ub = h_counts[known];
            if (ub <= 0) break;
            const int* n_dev = counts.p + 2 * it;
            int* queue = counts.p + 2 * it + 1;
            int* n_out = counts.p + 2 * (it + 1);
            if (pairs) {
                // at most ub pairs; ≈ ub/2 unless many regions are down to one active start — the grid covers 5/8·ub, warps stride
                const unsigned pgrid = (unsigned)(((int64_t)ub * 5 / 8 + E_WARPS - 1) / E_WARPS + 1);
                { KTimer kt(ctx, CPN_K_PREP);
                  k_make_pairs<<<blocks_for(ub, 256), 256, 0, s>>>(ub, cur, n_init, rt.rdesc, phi.p, NI, pdesc.p, phil.p, NI, n_dev,
                                                                   pair_cnt.p + 4 * it); }
                { KTimer kt(ctx, CPN_K_ESTEP);
                  if (views) k_estep_pair<true><<<pgrid, E_WARPS * 32, EP_SMEM, s>>>(pair_cnt.p + 4 * it, pdesc.p, phil.p, NI, rt, NI, ES.p);
                  else if (pair_ws) k_estep_pair_ws<<<std::min<unsigned>(pgrid, (unsigned)(ctx->sm_count * ctx->epw_blocks)), E_WARPS * 32, EPW_SMEM, s>>>(
                                        pair_cnt.p + 4 * it, pdesc.p, phil.p, NI, rt, NI, ES.p);
                  else k_estep_pair<false><<<pgrid, E_WARPS * 32, EP_SMEM, s>>>(pair_cnt.p + 4 * it, pdesc.p, phil.p, NI, rt, NI, ES.p); }
            } else {
            { KTimer kt(ctx, CPN_K_PREP);
              k_make_desc<<<blocks_for(ub, 256), 256, 0, s>>>(ub, cur, n_init, rt.rdesc, phi.p, NI, idesc.p, phil.p, NI, n_dev); }
            { KTimer kt(ctx, CPN_K_ESTEP);
              if (views) k_estep<true><<<estep_grid(ctx, ub), E_WARPS * 32, 0, s>>>(ub, idesc.p, phil.p, NI, rt, NI, ES.p, n_dev);
              else k_estep<false><<<estep_grid(ctx, ub), E_WARPS * 32, 0, s>>>(ub, idesc.p, phil.p, NI, rt, NI, ES.p, n_dev); }
            }
            if (mstep_mode == CPN_MSTEP_ANALYTIC) {
                unsigned grid = std::min<unsigned>(blocks_for(ub, M_THREADS), (unsigned)(ctx->sm_count * ws_blocks_per_sm));
                { KTimer kt(ctx, CPN_K_MSTEP);
                  k_mstep_ws<<<grid, M_THREADS, 0, s>>>(ub, cur, n_init, rt, st, max_em_iters, amax, bmax, cmax, queue, alive.p, n_dev); }
                int nblk = (ub + CP_TILE - 1) / CP_TILE;
                { KTimer kt(ctx, CPN_K_PREP); k_compact_count<<<nblk, CP_THREADS, 0, s>>>(ub, alive.p, blk_cnt.p, n_dev); }
                { KTimer kt(ctx, CPN_K_PREP); k_compact_scan<<<1, 1024, 0, s>>>(nblk, blk_cnt.p, n_out); }
                { KTimer kt(ctx, CPN_K_PREP); k_compact_scatter<<<nblk, CP_THREADS, 0, s>>>(ub, alive.p, cur, blk_cnt.p, nxt, n_dev); }
            } else {
                KTimer kt(ctx, CPN_K_MSTEP);
                k_mstep<<<blocks_for(ub, M_THREADS), M_THREADS, 0, s>>>(ub, cur, n_init, rt, st, max_em_iters, amax, bmax, cmax, mstep_mode,
                                                                         nxt, n_out, n_dev);
            }
            CPN_CUDA(ctx, cudaMemcpyAsync((void*)(ctx->h_counts + it + 1), n_out, sizeof(int), cudaMemcpyDeviceToHost, s));
            CPN_CUDA(ctx, cudaEventRecord(ctx->it_ev[it + 1], s));
            std::swap(cur, nxt);
        }
        trace(ctx, "em-done");
#ifdef EP_STATS
        if (pairs) {
            CPN_CUDA(ctx, cpn_wait_stream(ctx, s));
            std::vector<int> h_pc(4 * max_em_iters), h_ct(2 * (max_em_iters + 1));
            unsigned long long h_st[8];
            cudaMemcpy(h_pc.data(), pair_cnt.p, 4 * max_em_iters * sizeof(int), cudaMemcpyDeviceToHost);
            cudaMemcpy(h_ct.data(), counts.p, h_ct.size() * sizeof(int), cudaMemcpyDeviceToHost);
            cudaMemcpyFromSymbol(h_st, ep_stats, sizeof(h_st));
            fprintf(stderr, "[ep_stats] blocks + %llu  - %llu  general %llu | singleton pairs %llu  full pairs %llu  slow %llu\n", h_st[0], h_st[1],
                    h_st[2], h_st[3], h_st[4], h_st[5]);
            for (int it = 0; it < max_em_iters; ++it) fprintf(stderr, "[ep_stats] it %d active %d pairs c>=0 %d c<0 %d\n", it, h_ct[2 * it], h_pc[4 * it], h_pc[4 * it + 1]);
        }
#endif
        DevBuf<double> work;
        CPN_CUDA(ctx, work.alloc(4, s));
        CPN_CUDA(ctx, cudaMemsetAsync(work.p, 0, 4 * sizeof(double), s));
        { KTimer kt(ctx, CPN_K_PREP); k_work_sum<<<blocks_for(NI, 256), 256, 0, s>>>(NI, n_init, rt.rdesc, st, work.p); }
        double h_work[4] = {0, 0, 0, 0};
        CPN_CUDA(ctx, cudaMemcpyAsync(h_work, work.p, sizeof(h_work), cudaMemcpyDeviceToHost, s));     // read after the final wait below
        // score converged instances (em_algorithm.jl:344), then pick per region
        { KTimer kt(ctx, CPN_K_SCORE);
          k_score<<<blocks_for(NI, E_WARPS), E_WARPS * 32, 0, s>>>((int)NI, nullptr, n_init, rt, phi.p, NI, stat.p, 1, Qs.p); }
        DevOut<double> o_phi, o_Q, o_ps, o_Qs;
        DevOut<int32_t> o_status;
        DevOut<int64_t> o_cnt;
        CPN_CUDA(ctx, o_phi.bind(phi_hat, 3 * U, on_device, s));
        CPN_CUDA(ctx, o_Q.bind(Q, U, on_device, s));
        CPN_CUDA(ctx, o_status.bind(status, U, on_device, s));
        CPN_CUDA(ctx, o_cnt.bind(counters, 4 * U, on_device, s));
        if (phi_starts) CPN_CUDA(ctx, o_ps.bind(phi_starts, 3 * NI, on_device, s));
        if (Q_starts) CPN_CUDA(ctx, o_Qs.bind(Q_starts, NI, on_device, s));
        { KTimer kt(ctx, CPN_K_PICK);
          k_pick<<<blocks_for(U, 128), 128, 0, s>>>(U, n_init, st, Qs.p, o_phi.p, o_Q.p, o_status.p, o_cnt.p, phi_starts ? o_ps.p : nullptr,
                                                    Q_starts ? o_Qs.p : nullptr); }
        CPN_CUDA(ctx, cudaGetLastError());
        CPN_CUDA(ctx, o_phi.finish(s)); CPN_CUDA(ctx, o_Q.finish(s)); CPN_CUDA(ctx, o_status.finish(s)); CPN_CUDA(ctx, o_cnt.finish(s));
        if (phi_starts) CPN_CUDA(ctx, o_ps.finish(s));
        if (Q_starts) CPN_CUDA(ctx, o_Qs.finish(s));
        if (sa) {
            // rscle_grp_mod! + get_stat_sums! (nanopore.jl:338-343) on ϕ̂ straight from device memory; regions without a
            // fit carry NaN ϕ̂ and produce NaN summaries (the host maps them to proc=false).
            std::vector<int64_t> h_cpg(R + 1), h_sub(R + 1);
            if (offs_dev) {
                CPN_CUDA(ctx, cudaMemcpyAsync(h_cpg.data(), sa->cpg_off, (R + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
                CPN_CUDA(ctx, cudaMemcpyAsync(h_sub.data(), sa->sub_off, (R + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
                CPN_CUDA(ctx, cpn_wait_stream(ctx, s));
            } else {
                std::memcpy(h_cpg.data(), sa->cpg_off, (R + 1) * sizeof(int64_t));
                std::memcpy(h_sub.data(), sa->sub_off, (R + 1) * sizeof(int64_t));
            }
            int64_t sumN = h_cpg[R], sumK = h_sub[R], Nmax = 0;
            for (int64_t r = 0; r < R; ++r) {
                CPN_ARG(ctx, h_cpg[r + 1] - h_cpg[r] >= 2, "every region needs N >= 2 CpG sites");
                Nmax = std::max(Nmax, h_cpg[r + 1] - h_cpg[r]);
            }
            DevIn<int64_t> i_cpg, i_sub, i_sp, i_sq;
            DevIn<double> i_rho, i_dn;
            CPN_CUDA(ctx, i_cpg.bind(sa->cpg_off, R + 1, offs_dev, s)); CPN_CUDA(ctx, i_sub.bind(sa->sub_off, R + 1, offs_dev, s));
            CPN_CUDA(ctx, i_sp.bind(sa->sub_p, sumK, on_device, s)); CPN_CUDA(ctx, i_sq.bind(sa->sub_q, sumK, on_device, s));
            CPN_CUDA(ctx, i_rho.bind(sa->rho_n, sumN, on_device, s)); CPN_CUDA(ctx, i_dn.bind(sa->d_n, sumN - R, on_device, s));
            DevOut<double> s_logZ, s_ex, s_exx, s_lg, s_e1, s_e2, s_mml, s_nme;
            CPN_CUDA(ctx, s_logZ.bind(sa->logZ, R, on_device, s)); CPN_CUDA(ctx, s_ex.bind(sa->ex, sumN, on_device, s));
            CPN_CUDA(ctx, s_exx.bind(sa->exx, sumN - R, on_device, s)); CPN_CUDA(ctx, s_lg.bind(sa->log_g, 4 * sumK, on_device, s));
            CPN_CUDA(ctx, s_e1.bind(sa->e_log_g1, sumK, on_device, s)); CPN_CUDA(ctx, s_e2.bind(sa->e_log_g2, sumK, on_device, s));
            CPN_CUDA(ctx, s_mml.bind(sa->mml, sumK, on_device, s)); CPN_CUDA(ctx, s_nme.bind(sa->nme, sumK, on_device, s));
            int rc2 = cpn_stat_sums_launch(ctx, R, o_phi.p, nullptr, i_cpg.p, i_sub.p, i_cpg.p, i_rho.p, i_dn.p, i_sub.p, i_sp.p, i_sq.p, Nmax,
                                           s_logZ.p, s_ex.p, s_exx.p, s_lg.p, s_e1.p, s_e2.p, s_mml.p, s_nme.p, nullptr);
            if (rc2) return rc2;
            CPN_CUDA(ctx, s_logZ.finish(s)); CPN_CUDA(ctx, s_ex.finish(s)); CPN_CUDA(ctx, s_exx.finish(s)); CPN_CUDA(ctx, s_lg.finish(s));
            CPN_CUDA(ctx, s_e1.finish(s)); CPN_CUDA(ctx, s_e2.finish(s)); CPN_CUDA(ctx, s_mml.finish(s)); CPN_CUDA(ctx, s_nme.finish(s));
            CPN_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
            CPN_CUDA(ctx, cpn_wait_stream(ctx, s));
        } else {
            CPN_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
            CPN_CUDA(ctx, cpn_wait_stream(ctx, s));
        }
        for (int k = 0; k < 4; ++k) ctx->last_work[k] += h_work[k];
    }
    trace(ctx, "end");
    if (!pre) {
        float ms = 0;
        CPN_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->last_ms = ms;
    }
    if (ctx->profile) cpn_profile_collect(ctx);
    return CPN_OK;
}


// Null statistics of the two-sample test (two_samp_perm_tests.jl:65-86): for permutation i (models 2i and 2i+1) and
// sub-region k: Tmml = mml1 − mml2, Tnme = nme1 − nme2, Tcmd; all NaN when either refit found no optimum.
__global__ void k_two_samp_T(int64_t P, const int64_t* __restrict__ k_off /*[2P+1] per model*/, const int64_t* __restrict__ cmd_off /*[P+1]*/,
                             const double* __restrict__ mml, const double* __restrict__ nme, const double* __restrict__ cmd,
                             const int32_t* __restrict__ status, double* __restrict__ T) {
    int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (i >= P) return;
    const int64_t c0 = cmd_off[i], K = cmd_off[i + 1] - c0, ka = k_off[2 * i], kb = k_off[2 * i + 1];
    const bool bad = status[2 * i] < 0 || status[2 * i + 1] < 0;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int64_t k = lane; k < K; k += 32) {
        T[3 * c0 + k] = bad ? nan : mml[ka + k] - mml[kb + k];
        T[3 * c0 + K + k] = bad ? nan : nme[ka + k] - nme[kb + k];
        T[3 * c0 + 2 * K + k] = bad ? nan : cmd[c0 + k];
    }
}


// ---- two-sample test: views of the pooled reads, built on the device --------------------------------------------------------
// One entry per region of a chunk: where its views, read-index rows and summary tables start, and its sizes.
struct TsRegion { int64_t v0, row0, ex0, k0, c0, pid0, gr; int32_t P, n, n1, N, K, rows1, rows2, pad; };

// Thread = (region blockIdx.y, permutation i).  Splits the region's n pooled reads into sample 1 (perm_ids' ascending 1-based
// indices, two_samp_perm_tests.jl:50-59; or — perm_ids == null — the selection-sampling draw of cpn_philox.h) and the rest, in
// order, and writes the two views (descriptor, read-index rows padded with read 0, RNG key) plus the index tables the summaries
// and CMD kernels take for these 2·ΣP models.
__global__ void k_two_samp_views(const TsRegion* __restrict__ regs, const RegionDesc* __restrict__ base, const int32_t* __restrict__ perm_ids,
                                 uint64_t seed, int64_t V, int64_t Pt, int64_t sN, int64_t sK, int64_t sC, RegionDesc* __restrict__ vdesc,
                                 int32_t* __restrict__ ridx, int2* __restrict__ vkey, int64_t* __restrict__ reg_of, int64_t* __restrict__ ex_off,
                                 int64_t* __restrict__ k_off, int64_t* __restrict__ cmd_off, int64_t* __restrict__ pa, int64_t* __restrict__ pb) {
    const int r = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0 && i == 0) { ex_off[V] = sN; k_off[V] = sK; cmd_off[Pt] = sC; }
    const TsRegion t = regs[r];
    if (i >= t.P) return;
    const RegionDesc b = base[r];
    const int64_t v = t.v0 + 2 * (int64_t)i;
    const int64_t row = t.row0 + (int64_t)i * (t.rows1 + t.rows2);
    int32_t* r1 = ridx + row * 32;
    int32_t* r2 = r1 + (int64_t)t.rows1 * 32;
    int c1 = 0, c2 = 0;
    if (perm_ids) {
        const int32_t* ids = perm_ids + t.pid0 + (int64_t)i * t.n1;
        int j = 0;
        for (int q = 0; q < t.n; ++q) {
            const bool in1 = j < t.n1 && ids[j] - 1 == q;
            if (in1) { r1[c1++] = q; ++j; } else r2[c2++] = q;
        }
    } else {
        int need = t.n1;
        uint32_t x0 = 0, x1 = 0, x2 = 0, x3 = 0;
        for (int q = 0; q < t.n; ++q) {
            if ((q & 3) == 0) {
                const CpnPhilox4 blk = cpn_philox4x32_10(seed, (uint32_t)t.gr, (uint32_t)i, (uint32_t)(q >> 2), 1u);
                x0 = blk.x[0]; x1 = blk.x[1]; x2 = blk.x[2]; x3 = blk.x[3];
            }
            const uint32_t x = (q & 2) ? ((q & 1) ? x3 : x2) : ((q & 1) ? x1 : x0);
            const bool in1 = (int)cpn_mulhi32(x, (uint32_t)(t.n - q)) < need;
            if (in1) { r1[c1++] = q; --need; } else r2[c2++] = q;
        }
    }
    for (; c1 < t.rows1 * 32; ++c1) r1[c1] = 0;
    for (; c2 < t.rows2 * 32; ++c2) r2[c2] = 0;
    RegionDesc d;
    d.g0 = b.g0; d.ew_off = b.ew_off; d.rb_off = b.rb_off; d.L = b.L; d.pad = 0; d.llrmax = b.llrmax; d.idmax = b.idmax;
    d.M = t.n1; d.ridx = (int32_t)row;
    vdesc[v] = d;
    d.M = t.n - t.n1; d.ridx = (int32_t)(row + t.rows1);
    vdesc[v + 1] = d;
    vkey[v] = make_int2((int)t.gr, 2 * i); vkey[v + 1] = make_int2((int)t.gr, 2 * i + 1);
    reg_of[v] = r; reg_of[v + 1] = r;
    ex_off[v] = t.ex0 + 2 * (int64_t)i * t.N; ex_off[v + 1] = t.ex0 + (2 * (int64_t)i + 1) * t.N;
    k_off[v] = t.k0 + 2 * (int64_t)i * t.K; k_off[v + 1] = t.k0 + (2 * (int64_t)i + 1) * t.K;
    const int64_t pi = (t.v0 >> 1) + i;
    pa[pi] = v; pb[pi] = v + 1; cmd_off[pi] = t.c0 + (int64_t)i * t.K;
}

// Observed statistics of the two-sample test (two_samp_perm_tests.jl:165-167) from the summaries of the two fitted models
// (models 2r and 2r+1) and their CMD, in the [region][statistic][K] layout the counting kernel reads.
__global__ void k_two_samp_Tobs(int64_t R, const int64_t* __restrict__ sub_off, const double* __restrict__ mml, const double* __restrict__ nme,
                                const double* __restrict__ cmd, double* __restrict__ T) {
    int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (r >= R) return;
    const int64_t k0 = sub_off[r], K = sub_off[r + 1] - k0;
    for (int64_t k = lane; k < K; k += 32) {
        T[3 * k0 + k] = mml[2 * k0 + k] - mml[2 * k0 + K + k];
        T[3 * k0 + K + k] = nme[2 * k0 + k] - nme[2 * k0 + K + k];
        T[3 * k0 + 2 * K + k] = cmd[k0 + k];
    }
}

// Observed statistics and, when null statistics are given, the permutation p-values of Rc regions (host outputs).
static int two_samp_observed(cpn_ctx* ctx, int64_t Rc, const double* phi_s1, const double* phi_s2, const int64_t* po, const int64_t* tn,
                             const int32_t* exact, const int64_t* cp, const int64_t* so, int Kmax, int64_t Nmax, const int64_t* d_cpg,
                             const double* d_rho, const double* d_dn, const int64_t* d_so, const int64_t* d_sp, const int64_t* d_sq,
                             const double* d_Tnull, double* T_obs, int64_t* cnt, int64_t* n_valid, double* pval, bool ordered) {
    cudaStream_t s = ctx->stream;
    const int64_t T2 = 2 * Rc, sumN = cp[Rc], sumK = so[Rc];
    std::vector<double> h_phi(3 * T2);
    std::vector<int64_t> h_reg(T2), h_exo(T2 + 1), h_ko(T2 + 1), h_pa(Rc), h_pb(Rc), h_nul(Rc);
    std::vector<int32_t> h_exact(Rc, 0);
    for (int64_t r = 0; r < Rc; ++r) {
        for (int k = 0; k < 3; ++k) { h_phi[6 * r + k] = phi_s1[3 * r + k]; h_phi[6 * r + 3 + k] = phi_s2[3 * r + k]; }
        const int64_t N = cp[r + 1] - cp[r], K = so[r + 1] - so[r];
        h_reg[2 * r] = h_reg[2 * r + 1] = r;
        h_exo[2 * r] = 2 * cp[r]; h_exo[2 * r + 1] = 2 * cp[r] + N;
        h_ko[2 * r] = 2 * so[r]; h_ko[2 * r + 1] = 2 * so[r] + K;
        h_pa[r] = 2 * r; h_pb[r] = 2 * r + 1;
        h_nul[r] = tn[r] - tn[0];
        if (exact) h_exact[r] = exact[r];
    }
    h_exo[T2] = 2 * sumN; h_ko[T2] = 2 * sumK;
    DevBuf<double> d_phi, logZ, lg, e1, e2, mml, nme, cmd, d_T, o_p;
    DevBuf<int64_t> d_reg, d_exo, d_ko, d_pa, d_pb, d_po, d_nul, o_cnt, o_nv;
    DevBuf<int32_t> d_exact;
    CpnArenaScope scope(&ctx->stage_arena);
    CPN_CUDA(ctx, d_phi.upload(h_phi.data(), 3 * T2, s)); CPN_CUDA(ctx, d_reg.upload(h_reg.data(), T2, s));
    CPN_CUDA(ctx, d_exo.upload(h_exo.data(), T2 + 1, s)); CPN_CUDA(ctx, d_ko.upload(h_ko.data(), T2 + 1, s));
    CPN_CUDA(ctx, d_pa.upload(h_pa.data(), Rc, s)); CPN_CUDA(ctx, d_pb.upload(h_pb.data(), Rc, s));
    CPN_CUDA(ctx, d_po.upload(po, Rc + 1, s)); CPN_CUDA(ctx, d_nul.upload(h_nul.data(), Rc, s)); CPN_CUDA(ctx, d_exact.upload(h_exact.data(), Rc, s));
    CPN_CUDA(ctx, logZ.alloc(T2, s)); CPN_CUDA(ctx, lg.alloc(8 * sumK, s));
    CPN_CUDA(ctx, e1.alloc(2 * sumK, s)); CPN_CUDA(ctx, e2.alloc(2 * sumK, s)); CPN_CUDA(ctx, mml.alloc(2 * sumK, s)); CPN_CUDA(ctx, nme.alloc(2 * sumK, s));
    CPN_CUDA(ctx, cmd.alloc(sumK, s)); CPN_CUDA(ctx, d_T.alloc(3 * sumK, s));
    CPN_CUDA(ctx, o_cnt.alloc(3 * sumK, s)); CPN_CUDA(ctx, o_nv.alloc(3 * sumK, s)); CPN_CUDA(ctx, o_p.alloc(3 * sumK, s));
    int rc = cpn_summaries_and_cmd(ctx, ordered, T2, d_phi.p, d_reg.p, d_exo.p, d_ko.p, 2 * sumN, 2 * sumK, Rc, sumN, d_cpg, d_rho, d_dn, d_so, d_sp, d_sq,
                                   Nmax, Rc, d_pa.p, d_pb.p, d_so, logZ.p, lg.p, e1.p, e2.p, mml.p, nme.p, cmd.p);
    if (rc) return rc;
    { KTimer kt(ctx, CPN_K_SWEEP); k_two_samp_Tobs<<<blocks_for(Rc * 32, 256), 256, 0, s>>>(Rc, d_so, mml.p, nme.p, cmd.p, d_T.p); }
    CPN_CUDA(ctx, cudaGetLastError());
    // with no null statistics (P = 0 everywhere) the counting kernel still runs: zero counts, NaN p-values (n_valid ≤ 20)
    rc = cpn_two_samp_pvals_launch(ctx, Rc, Kmax, d_so, d_po.p, d_nul.p, d_exact.p, d_T.p, d_Tnull ? d_Tnull : d_T.p, o_cnt.p, o_nv.p, o_p.p);
    if (rc) return rc;
    CPN_CUDA(ctx, d_T.download(T_obs, 3 * sumK, s)); CPN_CUDA(ctx, o_cnt.download(cnt, 3 * sumK, s));
    CPN_CUDA(ctx, o_nv.download(n_valid, 3 * sumK, s)); CPN_CUDA(ctx, o_p.download(pval, 3 * sumK, s));
    CPN_CUDA(ctx, cpn_wait_stream(ctx, s));                       // host staging vectors and the DevBufs go out of scope
    return CPN_OK;
}

}  // namespace

extern "C" {

// Host-pointer entry points split a large batch into contiguous chunks of regions (balanced by M·L) and run them on worker
// threads, each with its own stream: the host→device copy of chunk k+1 overlaps the EM kernels of chunk k, and the tail of
// one chunk's kernels overlaps the head of the other's.  Regions are independent, so results do not depend on the split.

static int fit_chunked(cpn_ctx* ctx, bool on_device, int64_t R, const int64_t* grp_off_in, const double* Nl, const double* rho_l,
                       const double* d_l, const int64_t* read_off_in, const double* lu, const double* lm, const double* phi0, int32_t n_init,
                       int32_t max_em_iters, double amax, double bmax, double cmax, int32_t mstep_mode, double* phi_hat, double* Q,
                       int32_t* status, int64_t* counters, double* phi_starts, double* Q_starts, const SummaryArgs* sa_in, bool packed = false) {
    // Chunk plan: K chunks whose costs (M·L summed over regions) grow geometrically, so the first upload — the only one that
    // nothing overlaps — is short, while later chunks are big enough to keep the EM kernels' late, sparsely populated
    // iterations amortised.  Device-resident inputs have nothing to upload and run as one batch unless CPN_DEVICE_CHUNKS asks
    // for equal parts (kept for experiments and tests).
    // CPN_WORKERS=0 disables chunking; CPN_CHUNKS / CPN_CHUNK_RATIO / CPN_CHUNK_REGIONS tune the plan.
    const int n_workers = std::max(0, env_int("CPN_WORKERS", 2));
    const int64_t chunk_regions = std::max(1, env_int("CPN_CHUNK_REGIONS", 12500));
    // no chunk holds more than CPN_MAX_CHUNK_REGIONS regions: the staging arenas (one per chunk in flight, ≈ 33 KB per region)
    // stay bounded however large the batch is — a whole genome (3·10⁶ regions) runs as 12 chunks through the same five arenas
    const int64_t max_chunk_regions = std::max(1, env_int("CPN_MAX_CHUNK_REGIONS", 250000));
    int C = (int)((R > 0 && grp_off_in && read_off_in && n_workers >= 1) ? std::min<int64_t>(R / chunk_regions, 4) : 0);
    if (on_device) C = std::min(C, 1);     // measured: concurrent chunks do not beat one full-width batch (97.3 vs 97.1 ms per 1e5 regions)
    const char* cvar = on_device ? "CPN_DEVICE_CHUNKS" : "CPN_CHUNKS";
    if (R / chunk_regions > 1 && n_workers >= 1 && getenv(cvar)) C = (int)std::min<int64_t>(std::max(1, env_int(cvar, C)), R);
    if (C <= 1)
        return fit_impl(ctx, on_device, R, grp_off_in, Nl, rho_l, d_l, read_off_in, lu, lm, phi0, n_init, max_em_iters, amax, bmax, cmax,
                        mstep_mode, phi_hat, Q, status, counters, phi_starts, Q_starts, sa_in, nullptr, -1, packed);
    CPN_ARG(ctx, Nl && rho_l && d_l && lu && lm && phi_hat && Q, "null pointer");
    CPN_ARG(ctx, n_init >= 1, "n_init >= 1");
    CPN_CUDA(ctx, cudaSetDevice(ctx->device));
    auto t0 = std::chrono::steady_clock::now();
    // the CSR offsets are needed on the host to cut the batch; device-resident ones are fetched once
    std::vector<int64_t> h_off[4];
    const int64_t* grp_off = grp_off_in;
    const int64_t* read_off = read_off_in;
    SummaryArgs sa_host{};
    const SummaryArgs* sa = sa_in;
    if (on_device) {
        CPN_CUDA(ctx, cudaSetDevice(ctx->device));
        const int64_t* src[4] = {grp_off_in, read_off_in, sa_in ? sa_in->cpg_off : nullptr, sa_in ? sa_in->sub_off : nullptr};
        for (int k = 0; k < 4; ++k)
            if (src[k]) {
                h_off[k].resize(R + 1);
                CPN_CUDA(ctx, cudaMemcpyAsync(h_off[k].data(), src[k], (R + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
            }
        CPN_CUDA(ctx, cpn_wait_stream(ctx, ctx->stream));
        grp_off = h_off[0].data(); read_off = h_off[1].data();
        if (sa_in) { sa_host = *sa_in; sa_host.cpg_off = h_off[2].data(); sa_host.sub_off = h_off[3].data(); sa = &sa_host; }
    }
    const double ratio = on_device ? 1.0 : (getenv("CPN_CHUNK_RATIO") ? std::max(1.0, atof(getenv("CPN_CHUNK_RATIO"))) : 1.3);
    std::vector<int64_t> obs(R + 1);       // emission offsets = cumulative cost
    obs[0] = 0;
    for (int64_t r = 0; r < R; ++r) obs[r + 1] = obs[r] + (grp_off[r + 1] - grp_off[r]) * (read_off[r + 1] - read_off[r]);
    std::vector<int64_t> bnd(C + 1, R);
    bnd[0] = 0;
    {
        double total = 0, acc = 0, w = 1;
        for (int c = 0; c < C; ++c, w *= ratio) total += w;
        w = 1;
        for (int c = 1; c < C; ++c, w *= ratio) {
            acc += w;
            int64_t target = (int64_t)((double)obs[R] * (acc / total));
            bnd[c] = std::lower_bound(obs.begin(), obs.end(), target) - obs.begin();
            bnd[c] = std::min(std::max(bnd[c], bnd[c - 1]), R);
        }
    }
    {   // enforce the region cap: the geometric plan is kept while it respects it; from the first chunk that would exceed it on,
        // chunks of exactly the cap follow until the batch is covered
        bool over = false;
        for (int c = 0; c < C; ++c) over = over || (bnd[c + 1] - bnd[c] > max_chunk_regions);
        if (over) {
            std::vector<int64_t> nb{0};
            for (int c = 0; c < C && nb.back() < R; ++c) {
                if (bnd[c + 1] - nb.back() > max_chunk_regions) break;
                if (bnd[c + 1] > nb.back()) nb.push_back(bnd[c + 1]);
            }
            while (nb.back() < R) nb.push_back(std::min(R, nb.back() + max_chunk_regions));
            bnd = nb;
            C = (int)bnd.size() - 1;
        }
    }
    // children[0] stages chunks (host→device copies and the packing kernels) ahead of the W compute workers children[1..W]
    const int W = std::min(n_workers, C);
    while ((int)ctx->children.size() < W + 1) {
        cpn_ctx* ch = nullptr;
        if (cpn_create(ctx->device, &ch) != CPN_OK) { ctx->err = "could not create a worker context"; return CPN_ERR_CUDA; }
        if (ctx->children.empty() && env_int("CPN_UPLOAD_PRIO", 1)) {
            // the staging stream outranks the compute streams: its packing kernels take SM slots as soon as EM blocks retire
            int lo = 0, hi = 0;
            cudaStream_t ps = nullptr;
            if (cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess &&
                cudaStreamCreateWithPriority(&ps, cudaStreamNonBlocking, hi) == cudaSuccess) {
                cudaStreamDestroy(ch->stream);
                ch->stream = ps;
            }
        }
        ctx->children.push_back(ch);
    }
    for (int i = 0; i <= W; ++i) cpn_set_profile(ctx->children[i], ctx->profile ? 1 : 0);
    // device-side bracket of the whole call: ev0 precedes every worker stream's work, ev1 follows all of it
    CPN_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    for (int i = 0; i <= W; ++i) CPN_CUDA(ctx, cudaStreamWaitEvent(ctx->children[i]->stream, ctx->ev0, 0));
    const int depth = 2;                   // chunks staged ahead of the ones being computed
    while ((int)ctx->slot_arenas.size() < depth + W + 1) ctx->slot_arenas.push_back(new CpnArena());
    std::vector<CpnArena*> free_arenas(ctx->slot_arenas.begin(), ctx->slot_arenas.begin() + depth + W + 1);
    std::vector<std::unique_ptr<StagedBatch>> slot(C);
    std::vector<int> state(C, 0);          // 0 pending, 1 staged, -1 failed
    std::vector<int> rc(W + 1, CPN_OK);
    std::vector<int64_t> launches(W + 1, 0);
    std::vector<std::array<double, 4>> work(W + 1, std::array<double, 4>{0, 0, 0, 0});
    std::mutex mu;
    std::condition_variable cv;
    int taken = 0;
    bool abort_all = false;
    std::vector<std::thread> th;
    th.emplace_back([&]() {
        cpn_ctx* u = ctx->children[0];
        cudaSetDevice(ctx->device);                    // a new thread starts on device 0
        u->last_launches = 0;
        std::vector<int64_t> g, rd;
        for (int c = 0; c < C; ++c) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return abort_all || (c < taken + depth && !free_arenas.empty()); });
                if (abort_all) return;
            }
            const int64_t r0 = bnd[c], Rc = bnd[c + 1] - r0;
            int r = CPN_OK;
            if (Rc > 0) {
                CpnArena* ar;
                { std::lock_guard<std::mutex> lk(mu); ar = free_arenas.back(); free_arenas.pop_back(); }
                ar->reset();
                g.resize(Rc + 1); rd.resize(Rc + 1);
                for (int64_t i = 0; i <= Rc; ++i) { g[i] = grp_off[r0 + i] - grp_off[r0]; rd[i] = read_off[r0 + i] - read_off[r0]; }
                const int64_t g0 = grp_off[r0];
                slot[c].reset(new StagedBatch());
                slot[c]->arena = ar;
                slot[c]->seed = ctx->seed;
                slot[c]->unit_base = ctx->region_base + r0;   // seeded starts are keyed by the region's position in the whole batch
                trace(u, "stage");
                r = fit_stage(u, *slot[c], on_device, Rc, g.data(), Nl + g0, rho_l + g0, d_l + (g0 - r0), rd.data(), lu + obs[r0],
                              lm + (packed ? read_off[r0] : obs[r0]), phi0 ? phi0 + r0 * n_init * 3 : nullptr, n_init, 0, -1, packed);
                if (r == CPN_OK && cpn_wait_stream(u, u->stream) != cudaSuccess) { u->err = "stream synchronize failed"; r = CPN_ERR_CUDA; }
                launches[0] = u->last_launches;
                trace(u, "staged");
            }
            std::lock_guard<std::mutex> lk(mu);
            state[c] = r == CPN_OK ? 1 : -1;
            if (r != CPN_OK) { rc[0] = r; abort_all = true; }
            cv.notify_all();
            if (r != CPN_OK) return;
        }
    });
    for (int wi = 1; wi <= W; ++wi) {
        th.emplace_back([&, wi]() {
            cpn_ctx* w = ctx->children[wi];
            cudaSetDevice(ctx->device);
            std::vector<int64_t> cp, sb;
            for (;;) {
                int c;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    if (abort_all || taken >= C) return;
                    c = taken++;
                    cv.notify_all();
                    cv.wait(lk, [&] { return abort_all || state[c] != 0; });
                    if (state[c] != 1) return;
                }
                const int64_t r0 = bnd[c], Rc = bnd[c + 1] - r0;
                if (Rc <= 0) continue;
                SummaryArgs sc{};
                if (sa) {
                    cp.resize(Rc + 1); sb.resize(Rc + 1);
                    for (int64_t i = 0; i <= Rc; ++i) { cp[i] = sa->cpg_off[r0 + i] - sa->cpg_off[r0]; sb[i] = sa->sub_off[r0 + i] - sa->sub_off[r0]; }
                    const int64_t n0 = sa->cpg_off[r0], k0 = sa->sub_off[r0];
                    sc = SummaryArgs{cp.data(), sb.data(), sa->sub_p + k0, sa->sub_q + k0, sa->rho_n + n0, sa->d_n + (n0 - r0),
                                     sa->logZ + r0, sa->ex + n0, sa->exx + (n0 - r0), sa->log_g + 4 * k0, sa->e_log_g1 + k0, sa->e_log_g2 + k0,
                                     sa->mml + k0, sa->nme + k0};
                }
                int r = fit_impl(w, on_device, Rc, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, n_init, max_em_iters,
                                 amax, bmax, cmax, mstep_mode, phi_hat + 3 * r0, Q + r0, status ? status + r0 : nullptr,
                                 counters ? counters + 4 * r0 : nullptr, phi_starts ? phi_starts + r0 * n_init * 3 : nullptr,
                                 Q_starts ? Q_starts + r0 * n_init : nullptr, sa ? &sc : nullptr, slot[c].get(), 0);
                launches[wi] += w->last_launches;
                for (int k = 0; k < 4; ++k) work[wi][k] += w->last_work[k];
                // an error return may have left kernels that read the slot's tables in flight: drain the stream first
                if (r != CPN_OK) cudaStreamSynchronize(w->stream);
                {   // the worker's stream is idle: the staging slot can be reused
                    CpnArena* ar = slot[c]->arena;
                    slot[c].reset();
                    std::lock_guard<std::mutex> lk(mu);
                    free_arenas.push_back(ar);
                    cv.notify_all();
                }
                if (r != CPN_OK) {
                    std::lock_guard<std::mutex> lk(mu);
                    rc[wi] = r; abort_all = true;
                    cv.notify_all();
                    return;
                }
            }
        });
    }
    for (auto& t : th) t.join();
    slot.clear();
    ctx->last_launches = 0;
    for (double& w : ctx->last_work) w = 0;
    int ret = CPN_OK;
    for (int wi = 0; wi <= W; ++wi) {
        cpn_ctx* w = ctx->children[wi];
        ctx->last_launches += launches[wi];
        for (int k = 0; k < 4; ++k) { ctx->last_work[k] += work[wi][k]; }
        for (int k = 0; k < CPN_K_NCLASS; ++k) {
            ctx->prof_ms[k] += w->prof_ms[k]; ctx->prof_launches[k] += w->prof_launches[k];
            w->prof_ms[k] = 0; w->prof_launches[k] = 0;
        }
        if (rc[wi] != CPN_OK && ret == CPN_OK) { ret = rc[wi]; ctx->err = w->err; }
    }
    ctx->last_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (ret == CPN_OK) {
        for (int i = 0; i <= W; ++i) {
            CPN_CUDA(ctx, cudaEventRecord(ctx->children[i]->ev1, ctx->children[i]->stream));
            CPN_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->children[i]->ev1, 0));
        }
        CPN_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CPN_CUDA(ctx, cpn_wait_stream(ctx, ctx->stream));
        float ms = 0;
        CPN_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->last_ms = ms;
    }
    return ret;
}

// cpn_create_multi contexts: the regions of a host-pointer call are cut into one contiguous shard per device, balanced by
// reads x groups, and every device runs fit_chunked on its shard from its own host thread, writing into the caller's arrays at the
// shard's offset.  Per-region arithmetic does not depend on the shard (seeded starts are keyed by the global position).
static int multi_fit(cpn_ctx* ctx, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                     const int64_t* read_off, const double* lu, const double* lm, const double* phi0, int32_t n_init, int32_t max_em_iters,
                     double amax, double bmax, double cmax, int32_t mstep_mode, double* phi_hat, double* Q, int32_t* status, int64_t* counters,
                     double* phi_starts, double* Q_starts, const SummaryArgs* sa, bool packed = false) {
    CPN_ARG(ctx, R >= 0 && grp_off && read_off, "null pointer");
    const int nd = (int)ctx->devs.size();
    std::vector<int64_t> obs(R + 1);
    obs[0] = 0;
    for (int64_t r = 0; r < R; ++r) obs[r + 1] = obs[r] + (grp_off[r + 1] - grp_off[r]) * (read_off[r + 1] - read_off[r]);
    std::vector<int64_t> bnd(nd + 1, R);
    bnd[0] = 0;
    for (int d = 1; d < nd; ++d) {
        const int64_t target = (int64_t)((double)obs[R] * d / nd);
        bnd[d] = std::max<int64_t>(bnd[d - 1], std::lower_bound(obs.begin(), obs.end(), target) - obs.begin());
        bnd[d] = std::min(bnd[d], R);
    }
    std::vector<int> rc(nd, CPN_OK);
    std::vector<std::thread> th;
    for (int d = 0; d < nd; ++d) {
        th.emplace_back([&, d]() {
            cpn_ctx* c = ctx->devs[d];
            const int64_t r0 = bnd[d], Rd = bnd[d + 1] - r0;
            c->last_ms = 0; c->last_launches = 0;
            for (double& w : c->last_work) w = 0;
            if (Rd <= 0) return;
            c->seed = ctx->seed; c->region_base = ctx->region_base + r0;
            if (d > 0) cpn_set_profile(c, ctx->profile ? 1 : 0);
            std::vector<int64_t> g(Rd + 1), rd(Rd + 1), cp, sb;
            for (int64_t i = 0; i <= Rd; ++i) { g[i] = grp_off[r0 + i] - grp_off[r0]; rd[i] = read_off[r0 + i] - read_off[r0]; }
            const int64_t g0 = grp_off[r0];
            SummaryArgs sc{};
            if (sa) {
                cp.resize(Rd + 1); sb.resize(Rd + 1);
                for (int64_t i = 0; i <= Rd; ++i) { cp[i] = sa->cpg_off[r0 + i] - sa->cpg_off[r0]; sb[i] = sa->sub_off[r0 + i] - sa->sub_off[r0]; }
                const int64_t n0 = sa->cpg_off[r0], k0 = sa->sub_off[r0];
                sc = SummaryArgs{cp.data(), sb.data(), sa->sub_p + k0, sa->sub_q + k0, sa->rho_n + n0, sa->d_n + (n0 - r0), sa->logZ + r0, sa->ex + n0,
                                 sa->exx + (n0 - r0), sa->log_g + 4 * k0, sa->e_log_g1 + k0, sa->e_log_g2 + k0, sa->mml + k0, sa->nme + k0};
            }
            rc[d] = fit_chunked(c, false, Rd, g.data(), Nl + g0, rho_l + g0, d_l + (g0 - r0), rd.data(), lu + obs[r0],
                                lm + (packed ? read_off[r0] : obs[r0]),
                                phi0 ? phi0 + r0 * n_init * 3 : nullptr, n_init, max_em_iters, amax, bmax, cmax, mstep_mode, phi_hat + 3 * r0, Q + r0,
                                status ? status + r0 : nullptr, counters ? counters + 4 * r0 : nullptr,
                                phi_starts ? phi_starts + r0 * n_init * 3 : nullptr, Q_starts ? Q_starts + r0 * n_init : nullptr, sa ? &sc : nullptr, packed);
            c->region_base = 0;
        });
    }
    for (auto& t : th) t.join();
    cudaSetDevice(ctx->device);
    int ret = CPN_OK;
    double ms = 0, work[4] = {0, 0, 0, 0};
    int64_t launches = 0;
    for (int d = 0; d < nd; ++d) {
        cpn_ctx* c = ctx->devs[d];
        if (rc[d] != CPN_OK && ret == CPN_OK) { ret = rc[d]; if (d > 0) ctx->err = c->err; }
        ms = std::max(ms, c->last_ms); launches += c->last_launches;
        for (int k = 0; k < 4; ++k) work[k] += c->last_work[k];
        if (d > 0)
            for (int k = 0; k < CPN_K_NCLASS; ++k) {
                ctx->prof_ms[k] += c->prof_ms[k]; ctx->prof_launches[k] += c->prof_launches[k];
                c->prof_ms[k] = 0; c->prof_launches[k] = 0;
            }
    }
    ctx->last_ms = ms; ctx->last_launches = launches;
    for (int k = 0; k < 4; ++k) ctx->last_work[k] = work[k];
    return ret;
}

int cpn_fit_batch(cpn_ctx* ctx, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                  const int64_t* read_off, const double* log_pyx_u, const double* log_pyx_m, const double* phi0, int32_t n_init,
                  int32_t max_em_iters, double amax, double bmax, double cmax, int32_t mstep_mode, double* phi_hat, double* Q,
                  int32_t* status, int64_t* counters, double* phi_starts, double* Q_starts) {
    if (!ctx) return CPN_ERR_ARG;
    if (ctx->devs.size() > 1)
        return multi_fit(ctx, R, grp_off, Nl, rho_l, d_l, read_off, log_pyx_u, log_pyx_m, phi0, n_init, max_em_iters, amax, bmax, cmax, mstep_mode,
                         phi_hat, Q, status, counters, phi_starts, Q_starts, nullptr);
    return fit_chunked(ctx, false, R, grp_off, Nl, rho_l, d_l, read_off, log_pyx_u, log_pyx_m, phi0, n_init, max_em_iters, amax, bmax, cmax,
                            mstep_mode, phi_hat, Q, status, counters, phi_starts, Q_starts, nullptr);
}
int cpn_fit_batch_device(cpn_ctx* ctx, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                         const int64_t* read_off, const double* log_pyx_u, const double* log_pyx_m, const double* phi0, int32_t n_init,
                         int32_t max_em_iters, double amax, double bmax, double cmax, int32_t mstep_mode, double* phi_hat, double* Q,
                         int32_t* status, int64_t* counters, double* phi_starts, double* Q_starts) {
    if (!ctx) return CPN_ERR_ARG;
    return fit_chunked(ctx, true, R, grp_off, Nl, rho_l, d_l, read_off, log_pyx_u, log_pyx_m, phi0, n_init, max_em_iters, amax, bmax, cmax,
                       mstep_mode, phi_hat, Q, status, counters, phi_starts, Q_starts, nullptr);
}

static int analyze_common(cpn_ctx* ctx, bool dev, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                          const int64_t* read_off, const double* lu, const double* lm, const double* phi0, int32_t n_init, int32_t max_em_iters,
                          double amax, double bmax, double cmax, int32_t mstep_mode, const int64_t* cpg_off, const double* rho_n,
                          const double* d_n, const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q, double* phi_hat, double* Q,
                          int32_t* status, int64_t* counters, double* logZ, double* ex, double* exx, double* log_g, double* e_log_g1,
                          double* e_log_g2, double* mml, double* nme, bool packed = false) {
    if (!ctx) return CPN_ERR_ARG;
    CPN_ARG(ctx, cpg_off && rho_n && d_n && sub_off && sub_p && sub_q, "null per-CpG input pointer");
    SummaryArgs sa{cpg_off, sub_off, sub_p, sub_q, rho_n, d_n, logZ, ex, exx, log_g, e_log_g1, e_log_g2, mml, nme};
    if (!dev) {
        CPN_ARG(ctx, logZ && ex && exx && log_g && e_log_g1 && e_log_g2 && mml && nme, "null output pointer");
        if (ctx->devs.size() > 1)
            return multi_fit(ctx, R, grp_off, Nl, rho_l, d_l, read_off, lu, lm, phi0, n_init, max_em_iters, amax, bmax, cmax, mstep_mode, phi_hat, Q,
                             status, counters, nullptr, nullptr, &sa, packed);
        return fit_chunked(ctx, false, R, grp_off, Nl, rho_l, d_l, read_off, lu, lm, phi0, n_init, max_em_iters, amax, bmax, cmax, mstep_mode,
                                phi_hat, Q, status, counters, nullptr, nullptr, &sa, packed);
    }
    return fit_chunked(ctx, true, R, grp_off, Nl, rho_l, d_l, read_off, lu, lm, phi0, n_init, max_em_iters, amax, bmax, cmax, mstep_mode,
                       phi_hat, Q, status, counters, nullptr, nullptr, &sa, packed);
}

int cpn_analyze_batch(cpn_ctx* ctx, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                      const int64_t* read_off, const double* log_pyx_u, const double* log_pyx_m, const double* phi0, int32_t n_init,
                      int32_t max_em_iters, double amax, double bmax, double cmax, int32_t mstep_mode, const int64_t* cpg_off,
                      const double* rho_n, const double* d_n, const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q,
                      double* phi_hat, double* Q, int32_t* status, int64_t* counters, double* logZ, double* ex, double* exx, double* log_g,
                      double* e_log_g1, double* e_log_g2, double* mml, double* nme) {
    return analyze_common(ctx, false, R, grp_off, Nl, rho_l, d_l, read_off, log_pyx_u, log_pyx_m, phi0, n_init, max_em_iters, amax, bmax, cmax,
                          mstep_mode, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, phi_hat, Q, status, counters, logZ, ex, exx, log_g, e_log_g1,
                          e_log_g2, mml, nme);
}
int cpn_analyze_batch_device(cpn_ctx* ctx, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                             const int64_t* read_off, const double* log_pyx_u, const double* log_pyx_m, const double* phi0, int32_t n_init,
                             int32_t max_em_iters, double amax, double bmax, double cmax, int32_t mstep_mode, const int64_t* cpg_off,
                             const double* rho_n, const double* d_n, const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q,
                             double* phi_hat, double* Q, int32_t* status, int64_t* counters, double* logZ, double* ex, double* exx,
                             double* log_g, double* e_log_g1, double* e_log_g2, double* mml, double* nme) {
    return analyze_common(ctx, true, R, grp_off, Nl, rho_l, d_l, read_off, log_pyx_u, log_pyx_m, phi0, n_init, max_em_iters, amax, bmax, cmax,
                          mstep_mode, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, phi_hat, Q, status, counters, logZ, ex, exx, log_g, e_log_g1,
                          e_log_g2, mml, nme);
}

// Packed emissions (cpn_pack_emissions): one log-likelihood ratio per (read, group) and one Σ_obs log p(y|−1) per read — what
// the kernels consume, and half the bytes of the two log-likelihood tables on the way to the device.
int cpn_fit_batch_packed(cpn_ctx* ctx, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                         const int64_t* read_off, const double* llr, const double* read_base, const double* phi0, int32_t n_init,
                         int32_t max_em_iters, double amax, double bmax, double cmax, int32_t mstep_mode, double* phi_hat, double* Q,
                         int32_t* status, int64_t* counters, double* phi_starts, double* Q_starts) {
    if (!ctx) return CPN_ERR_ARG;
    if (ctx->devs.size() > 1)
        return multi_fit(ctx, R, grp_off, Nl, rho_l, d_l, read_off, llr, read_base, phi0, n_init, max_em_iters, amax, bmax, cmax, mstep_mode,
                         phi_hat, Q, status, counters, phi_starts, Q_starts, nullptr, true);
    return fit_chunked(ctx, false, R, grp_off, Nl, rho_l, d_l, read_off, llr, read_base, phi0, n_init, max_em_iters, amax, bmax, cmax, mstep_mode,
                       phi_hat, Q, status, counters, phi_starts, Q_starts, nullptr, true);
}
int cpn_analyze_batch_packed(cpn_ctx* ctx, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                             const int64_t* read_off, const double* llr, const double* read_base, const double* phi0, int32_t n_init,
                             int32_t max_em_iters, double amax, double bmax, double cmax, int32_t mstep_mode, const int64_t* cpg_off,
                             const double* rho_n, const double* d_n, const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q,
                             double* phi_hat, double* Q, int32_t* status, int64_t* counters, double* logZ, double* ex, double* exx,
                             double* log_g, double* e_log_g1, double* e_log_g2, double* mml, double* nme) {
    return analyze_common(ctx, false, R, grp_off, Nl, rho_l, d_l, read_off, llr, read_base, phi0, n_init, max_em_iters, amax, bmax, cmax,
                          mstep_mode, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, phi_hat, Q, status, counters, logZ, ex, exx, log_g, e_log_g1,
                          e_log_g2, mml, nme, true);
}

// Batched pmap_analyze_reg_bam (wgbs.jl:335-389): per region get_ϕhat_wgbs! (em_algorithm.jl:556-575: the same multi-start EM,
// em_iter_wgbs :412-470 being em_iter with the hard emissions of the calls) and get_stat_sums!.  WGBS regions are modelled at
// CpG resolution (get_grp_info! with min_grp_dist = 1, wgbs.jl:351: every CpG site is its own group, N_l = 1, ρ_l = ρ_n,
// d_l = d_n), so there is no rscle_grp_mod! step and the fit and the summaries share the per-CpG tables.
int cpn_analyze_wgbs_batch(cpn_ctx* ctx, int64_t R, const int64_t* cpg_off, const double* rho_n, const double* d_n, const int64_t* read_off,
                           const int8_t* calls, const double* phi0, int32_t n_init, int32_t max_em_iters, double amax, double bmax,
                           double cmax, int32_t mstep_mode, const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q,
                           double* phi_hat, double* Q, int32_t* status, int64_t* counters, double* logZ, double* ex, double* exx,
                           double* log_g, double* e_log_g1, double* e_log_g2, double* mml, double* nme) {
    if (!ctx) return CPN_ERR_ARG;
    CPN_ARG(ctx, R >= 0 && (R == 0 || (cpg_off && rho_n && d_n && read_off && calls)), "null pointer");
    if (R == 0) return CPN_OK;
    int64_t cells = 0;
    for (int64_t r = 0; r < R; ++r) cells += (cpg_off[r + 1] - cpg_off[r]) * (read_off[r + 1] - read_off[r]);
    std::vector<double> llr((size_t)cells), base((size_t)read_off[R]), ones((size_t)cpg_off[R], 1.0);
    int rc = cpn_wgbs_pack(R, cpg_off, read_off, calls, llr.data(), base.data(), 0);
    CPN_ARG(ctx, rc == CPN_OK, "calls must be -1, 0 or +1");
    return analyze_common(ctx, false, R, cpg_off, ones.data(), rho_n, d_n, read_off, llr.data(), base.data(), phi0, n_init, max_em_iters, amax, bmax,
                          cmax, mstep_mode, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, phi_hat, Q, status, counters, logZ, ex, exx, log_g,
                          e_log_g1, e_log_g2, mml, nme, true);
}

int cpn_estep_batch(cpn_ctx* ctx, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                    const int64_t* read_off, const double* log_pyx_u, const double* log_pyx_m, const double* phi, double* ES,
                    double* ave_logpy) {
    if (!ctx) return CPN_ERR_ARG;
    CPN_ARG(ctx, R >= 0 && grp_off && Nl && rho_l && d_l && read_off && log_pyx_u && log_pyx_m && phi && ES, "null pointer");
    CPN_CUDA(ctx, cudaSetDevice(ctx->device));
    LaunchCounter lc(ctx);
    if (R == 0) return CPN_OK;
    cudaStream_t s = ctx->stream;
    CPN_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    {
        PackedRegions pk;
        int rc = pack_regions(ctx, pk, false, true, R, grp_off, Nl, rho_l, d_l, read_off, log_pyx_u, log_pyx_m);
        if (rc) return rc;
        DevBuf<double> d_phi_in, d_phi, d_ES, d_ESo, d_Q;
        CPN_CUDA(ctx, d_phi_in.upload(phi, 3 * R, s));
        CPN_CUDA(ctx, d_phi.alloc(3 * R, s)); CPN_CUDA(ctx, d_ES.alloc(3 * R, s)); CPN_CUDA(ctx, d_ESo.alloc(3 * R, s));
        CPN_CUDA(ctx, d_Q.alloc(R, s));
        { KTimer kt(ctx, CPN_K_PREP); k_pack_phi<<<blocks_for(R, 256), 256, 0, s>>>(R, d_phi_in.p, d_phi.p); }
        DevBuf<ItemDesc> idesc;
        DevBuf<double> phil;
                CPN_CUDA(ctx, idesc.alloc(R, s)); CPN_CUDA(ctx, phil.alloc(3 * R, s));
        { KTimer kt(ctx, CPN_K_PREP);
          k_make_desc<<<blocks_for(R, 256), 256, 0, s>>>((int)R, nullptr, 1, pk.rt.rdesc, d_phi.p, R, idesc.p, phil.p, R, nullptr); }
        { KTimer kt(ctx, CPN_K_ESTEP); k_estep<false><<<estep_grid(ctx, R), E_WARPS * 32, 0, s>>>((int)R, idesc.p, phil.p, R, pk.rt, R, d_ES.p, nullptr); }
        { KTimer kt(ctx, CPN_K_PREP); k_unpack3<<<blocks_for(R, 256), 256, 0, s>>>(R, d_ES.p, d_ESo.p); }
        CPN_CUDA(ctx, d_ESo.download(ES, 3 * R, s));
        if (ave_logpy) {
            { KTimer kt(ctx, CPN_K_SCORE);
              k_score<<<blocks_for(R, E_WARPS), E_WARPS * 32, 0, s>>>((int)R, nullptr, 1, pk.rt, d_phi.p, R, nullptr, 0, d_Q.p); }
            CPN_CUDA(ctx, d_Q.download(ave_logpy, R, s));
        }
        CPN_CUDA(ctx, cudaGetLastError());
        CPN_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
        CPN_CUDA(ctx, cpn_wait_stream(ctx, s));
    }
    float ms = 0;
    CPN_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    if (ctx->profile) cpn_profile_collect(ctx);
    return CPN_OK;
}

// Test hook for the paired E-step: n_init parameter vectors per region (phi [R·n_init][3]), paired and evaluated exactly as one EM
// iteration of cpn_fit_batch does (k_make_pairs → k_estep_pair); ES [R·n_init][3].
int cpn_estep_batch_starts(cpn_ctx* ctx, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                           const int64_t* read_off, const double* log_pyx_u, const double* log_pyx_m, int32_t n_init, const double* phi,
                           double* ES) {
    if (!ctx) return CPN_ERR_ARG;
    CPN_ARG(ctx, R >= 0 && n_init >= 1 && grp_off && Nl && rho_l && d_l && read_off && log_pyx_u && log_pyx_m && phi && ES, "null pointer");
    CPN_ARG(ctx, R * (int64_t)n_init < (int64_t)2147483647, "R*n_init must fit in int32");
    CPN_CUDA(ctx, cudaSetDevice(ctx->device));
    LaunchCounter lc(ctx);
    if (R == 0) return CPN_OK;
    cudaStream_t s = ctx->stream;
    {
        PackedRegions pk;
        int rc = pack_regions(ctx, pk, false, true, R, grp_off, Nl, rho_l, d_l, read_off, log_pyx_u, log_pyx_m);
        if (rc) return rc;
        const int64_t NI = R * n_init;
        DevBuf<double> d_phi_in, d_phi, d_ES, d_ESo, phil;
        DevBuf<PairDesc> pdesc;
        DevBuf<int> pair_cnt;
        CPN_CUDA(ctx, d_phi_in.upload(phi, 3 * NI, s));
        CPN_CUDA(ctx, d_phi.alloc(3 * NI, s)); CPN_CUDA(ctx, d_ES.alloc(3 * NI, s)); CPN_CUDA(ctx, d_ESo.alloc(3 * NI, s));
        CPN_CUDA(ctx, pdesc.alloc(NI, s)); CPN_CUDA(ctx, phil.alloc(6 * NI, s)); CPN_CUDA(ctx, pair_cnt.alloc(4, s));
        CPN_CUDA(ctx, cudaMemsetAsync(pair_cnt.p, 0, 4 * sizeof(int), s));
        CPN_CUDA(ctx, cudaMemsetAsync(d_ES.p, 0xff, 3 * NI * sizeof(double), s));       // NaN: an instance no pair covers would show
        CPN_CUDA(ctx, ep_prepare(ctx));
        { KTimer kt(ctx, CPN_K_PREP); k_pack_phi<<<blocks_for(NI, 256), 256, 0, s>>>(NI, d_phi_in.p, d_phi.p); }
        { KTimer kt(ctx, CPN_K_PREP);
          k_make_pairs<<<blocks_for(NI, 256), 256, 0, s>>>((int)NI, nullptr, n_init, pk.rt.rdesc, d_phi.p, NI, pdesc.p, phil.p, NI, nullptr, pair_cnt.p); }
        const unsigned pgrid = (unsigned)((NI * 5 / 8 + E_WARPS - 1) / E_WARPS + 1);
        { KTimer kt(ctx, CPN_K_ESTEP);
          if (env_int("CPN_ESTEP_WS", 1) != 0)
              k_estep_pair_ws<<<std::min<unsigned>(pgrid, (unsigned)(ctx->sm_count * ctx->epw_blocks)), E_WARPS * 32, EPW_SMEM, s>>>(
                  pair_cnt.p, pdesc.p, phil.p, NI, pk.rt, NI, d_ES.p);
          else k_estep_pair<false><<<pgrid, E_WARPS * 32, EP_SMEM, s>>>(pair_cnt.p, pdesc.p, phil.p, NI, pk.rt, NI, d_ES.p); }
        { KTimer kt(ctx, CPN_K_PREP); k_unpack3<<<blocks_for(NI, 256), 256, 0, s>>>(NI, d_ES.p, d_ESo.p); }
        CPN_CUDA(ctx, d_ESo.download(ES, 3 * NI, s));
        CPN_CUDA(ctx, cudaGetLastError());
        CPN_CUDA(ctx, cpn_wait_stream(ctx, s));
    }
    if (ctx->profile) cpn_profile_collect(ctx);
    return CPN_OK;
}

int cpn_mstep_batch(cpn_ctx* ctx, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                    const int64_t* n_reads, const double* ES, const double* phi_init, int32_t mstep_mode, double* sol, int32_t* converged,
                    int64_t* counters) {
    if (!ctx) return CPN_ERR_ARG;
    CPN_ARG(ctx, R >= 0 && grp_off && Nl && rho_l && d_l && n_reads && ES && phi_init && sol && converged, "null pointer");
    CPN_ARG(ctx, mstep_mode == CPN_MSTEP_FD || mstep_mode == CPN_MSTEP_ANALYTIC, "mstep_mode");
    CPN_CUDA(ctx, cudaSetDevice(ctx->device));
    LaunchCounter lc(ctx);
    if (R == 0) return CPN_OK;
    cudaStream_t s = ctx->stream;
    CPN_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    {
        PackedRegions pk;
        int rc = pack_regions(ctx, pk, false, false, R, grp_off, Nl, rho_l, d_l, nullptr, nullptr, nullptr);
        if (rc) return rc;
        DevBuf<int64_t> d_nr, d_cnt;
        DevBuf<double> d_ES, d_phi, d_sol;
        DevBuf<int> d_conv;
        CPN_CUDA(ctx, d_nr.upload(n_reads, R, s)); CPN_CUDA(ctx, d_ES.upload(ES, 3 * R, s)); CPN_CUDA(ctx, d_phi.upload(phi_init, 3 * R, s));
        CPN_CUDA(ctx, d_sol.alloc(3 * R, s)); CPN_CUDA(ctx, d_conv.alloc(R, s)); CPN_CUDA(ctx, d_cnt.alloc(2 * R, s));
        { KTimer kt(ctx, CPN_K_MSTEP);
          k_mstep_once<<<blocks_for(R, 64), 64, 0, s>>>(R, pk.rt, d_nr.p, d_ES.p, d_phi.p, mstep_mode, d_sol.p, d_conv.p, d_cnt.p); }
        CPN_CUDA(ctx, cudaGetLastError());
        CPN_CUDA(ctx, d_sol.download(sol, 3 * R, s));
        CPN_CUDA(ctx, d_conv.download(converged, R, s));
        if (counters) CPN_CUDA(ctx, d_cnt.download(counters, 2 * R, s));
        CPN_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
        CPN_CUDA(ctx, cpn_wait_stream(ctx, s));
    }
    float ms = 0;
    CPN_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    if (ctx->profile) cpn_profile_collect(ctx);
    return CPN_OK;
}

// Two-sample permutation test for R regions, cut into chunks of regions whose refits (views) fit a bounded workspace.
// T_obs / cnt / n_valid / pval are produced when the two observed models are given; T_null and fit_status are optional.
static int two_samp_impl(cpn_ctx* ctx, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                         const int64_t* read_off, const double* lu, const double* lm, const int64_t* n1, const double* phi_s1,
                         const double* phi_s2, const int64_t* perm_off, const int32_t* perm_ids, const int32_t* exact, const double* phi0,
                         int32_t n_init, int32_t max_em_iters, double amax, double bmax, double cmax, int32_t mstep_mode,
                         const int64_t* cpg_off, const double* rho_n, const double* d_n, const int64_t* sub_off, const int64_t* sub_p,
                         const int64_t* sub_q, double* T_obs, int64_t* cnt, int64_t* n_valid, double* pval, double* T_null,
                         int32_t* fit_status) {
    CPN_ARG(ctx, R >= 0 && n_init >= 1 && max_em_iters >= 1, "R >= 0, n_init >= 1, max_em_iters >= 1");
    CPN_ARG(ctx, grp_off && Nl && rho_l && d_l && read_off && lu && lm && perm_off && n1, "null pointer");
    CPN_ARG(ctx, cpg_off && rho_n && d_n && sub_off && sub_p && sub_q, "null pointer");
    const bool observed = phi_s1 != nullptr;
    CPN_ARG(ctx, !observed || (phi_s2 && T_obs && cnt && n_valid && pval), "the test needs both observed models and the four output tables");
    CPN_ARG(ctx, observed || T_null, "nothing to compute: neither observed models nor a T_null output");
    CPN_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->last_launches = 0;
    ctx->last_ms = 0;
    if (R == 0) return CPN_OK;
    cudaStream_t s = ctx->stream;
    // ---- host: validation, cumulative tables, chunk plan ----------------------------------------------------------------
    std::vector<int64_t> obs(R + 1), pid(R + 1), tn(R + 1);     // emissions, perm_ids elements, T_null rows·K before region r
    obs[0] = pid[0] = tn[0] = 0;
    int64_t Nmax = 0, Kmax = 1;
    for (int64_t r = 0; r < R; ++r) {
        const int64_t L = grp_off[r + 1] - grp_off[r], n = read_off[r + 1] - read_off[r], P = perm_off[r + 1] - perm_off[r];
        const int64_t N = cpg_off[r + 1] - cpg_off[r], K = sub_off[r + 1] - sub_off[r];
        CPN_ARG(ctx, P >= 0, "perm_off must be non-decreasing");
        CPN_ARG(ctx, n1[r] >= 1 && n1[r] < n, "each sample needs at least one read");
        CPN_ARG(ctx, N >= 2, "every region needs N >= 2 CpG sites");
        Nmax = std::max(Nmax, N); Kmax = std::max(Kmax, K);
        obs[r + 1] = obs[r] + L * n; pid[r + 1] = pid[r] + P * n1[r]; tn[r + 1] = tn[r] + P * K;
    }
    if (perm_ids)
        for (int64_t r = 0; r < R; ++r) {
            const int64_t n = read_off[r + 1] - read_off[r], m1 = n1[r];
            const int32_t* ids = perm_ids + pid[r];
            for (int64_t i = 0; i < perm_off[r + 1] - perm_off[r]; ++i, ids += m1) {
                int32_t prev = 0;
                for (int64_t j = 0; j < m1; ++j) {
                    CPN_ARG(ctx, ids[j] > prev && ids[j] <= n, "perm_ids must be ascending 1-based read indices");
                    prev = ids[j];
                }
            }
        }
    const bool ordered = cpn_subregions_ordered(R, sub_off, sub_p, sub_q, cpg_off) && env_int("CPN_CMD_LANE", 1) != 0;
    const int64_t max_views = std::max<int64_t>(2, env_int("CPN_TWO_SAMP_MAX_VIEWS", 1 << 20));
    std::vector<int64_t> bnd{0};
    {
        int64_t v = 0;
        for (int64_t r = 0; r < R; ++r) {
            const int64_t vr = 2 * (perm_off[r + 1] - perm_off[r]);
            CPN_ARG(ctx, vr * (int64_t)n_init < (int64_t)2147483647, "2·P·n_init of one region must fit in int32");
            if (v > 0 && v + vr > max_views) { bnd.push_back(r); v = 0; }
            v += vr;
        }
        bnd.push_back(R);
    }
    CPN_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    int64_t launches = 0;
    double work_acc[4] = {0, 0, 0, 0};
    std::vector<int64_t> g, rd, cp, so, po;
    for (size_t c = 0; c + 1 < bnd.size(); ++c) {
        const int64_t r0 = bnd[c], Rc = bnd[c + 1] - r0;
        const int64_t Pt = perm_off[r0 + Rc] - perm_off[r0], V = 2 * Pt;
        g.resize(Rc + 1); rd.resize(Rc + 1); cp.resize(Rc + 1); so.resize(Rc + 1); po.resize(Rc + 1);
        for (int64_t i = 0; i <= Rc; ++i) {
            g[i] = grp_off[r0 + i] - grp_off[r0]; rd[i] = read_off[r0 + i] - read_off[r0]; cp[i] = cpg_off[r0 + i] - cpg_off[r0];
            so[i] = sub_off[r0 + i] - sub_off[r0]; po[i] = perm_off[r0 + i] - perm_off[r0];
        }
        const int64_t g0 = grp_off[r0], n0 = cpg_off[r0], k0 = sub_off[r0];
        const int64_t sumN = cp[Rc], sumK = so[Rc];
        // per-region layout of the chunk's views
        std::vector<TsRegion> h_regs(Rc);
        int64_t rows = 0, exo = 0, ko = 0, co = 0;
        int Pmax = 0;
        for (int64_t i = 0; i < Rc; ++i) {
            TsRegion& t = h_regs[i];
            const int64_t n = rd[i + 1] - rd[i], P = po[i + 1] - po[i];
            t.v0 = 2 * po[i]; t.row0 = rows; t.ex0 = exo; t.k0 = ko; t.c0 = co; t.pid0 = pid[r0 + i] - pid[r0]; t.gr = ctx->region_base + r0 + i;
            t.P = (int32_t)P; t.n = (int32_t)n; t.n1 = (int32_t)n1[r0 + i]; t.N = (int32_t)(cp[i + 1] - cp[i]); t.K = (int32_t)(so[i + 1] - so[i]);
            t.rows1 = (int32_t)((t.n1 + 31) / 32); t.rows2 = (int32_t)((n - t.n1 + 31) / 32); t.pad = 0;
            rows += P * (t.rows1 + t.rows2); exo += 2 * P * t.N; ko += 2 * P * t.K; co += P * t.K;
            Pmax = std::max<int>(Pmax, (int)P);
        }
        CPN_ARG(ctx, rows < (int64_t)2147483647, "too many read-index rows in one chunk (lower CPN_TWO_SAMP_MAX_VIEWS)");
        const int64_t sN = exo, sK = ko, sC = co;
        int rc = CPN_OK;
        if (V > 0) {
            CPN_CUDA(ctx, ctx->stage_arena.reset());
            StagedBatch sb;
            sb.arena = &ctx->stage_arena;
            sb.seed = ctx->seed;
            rc = fit_stage(ctx, sb, false, Rc, g.data(), Nl + g0, rho_l + g0, d_l + (g0 - r0), rd.data(), lu + obs[r0], lm + obs[r0],
                           phi0 ? phi0 + 2 * perm_off[r0] * n_init * 3 : nullptr, n_init, -1, V);
            if (rc) return rc;
            launches += ctx->last_launches;
            DevBuf<int64_t> d_reg, d_exo, d_ko, d_co, d_pa, d_pb, d_cpg, d_so, d_sp, d_sq;
            DevBuf<double> d_rho, d_dn, d_phi, d_Q, logZ, lg, e1, e2, mml, nme, cmd, T;
            DevBuf<int32_t> d_stat, d_pid;
            DevBuf<TsRegion> d_regs;
            {
                CpnArenaScope scope(&ctx->stage_arena);
                sb.V = V;
                sb.h_vL.resize(V);
                for (int64_t i = 0; i < Rc; ++i) std::fill(sb.h_vL.begin() + 2 * po[i], sb.h_vL.begin() + 2 * po[i + 1], (int32_t)(g[i + 1] - g[i]));
                CPN_CUDA(ctx, sb.vdesc.alloc(V, s)); CPN_CUDA(ctx, sb.ridx.alloc(rows * 32, s)); CPN_CUDA(ctx, sb.vkey.alloc(V, s));
                CPN_CUDA(ctx, d_regs.upload(h_regs.data(), Rc, s));
                if (perm_ids) CPN_CUDA(ctx, d_pid.upload(perm_ids + pid[r0], pid[r0 + Rc] - pid[r0], s));
                CPN_CUDA(ctx, d_reg.alloc(V, s)); CPN_CUDA(ctx, d_exo.alloc(V + 1, s)); CPN_CUDA(ctx, d_ko.alloc(V + 1, s));
                CPN_CUDA(ctx, d_co.alloc(Pt + 1, s)); CPN_CUDA(ctx, d_pa.alloc(Pt, s)); CPN_CUDA(ctx, d_pb.alloc(Pt, s));
                CPN_CUDA(ctx, d_cpg.upload(cp.data(), Rc + 1, s)); CPN_CUDA(ctx, d_so.upload(so.data(), Rc + 1, s));
                CPN_CUDA(ctx, d_sp.upload(sub_p + k0, sumK, s)); CPN_CUDA(ctx, d_sq.upload(sub_q + k0, sumK, s));
                CPN_CUDA(ctx, d_rho.upload(rho_n + n0, sumN, s)); CPN_CUDA(ctx, d_dn.upload(d_n + (n0 - r0), sumN - Rc, s));
                CPN_CUDA(ctx, d_phi.alloc(3 * V, s)); CPN_CUDA(ctx, d_Q.alloc(V, s)); CPN_CUDA(ctx, d_stat.alloc(V, s));
                CPN_CUDA(ctx, logZ.alloc(V, s)); CPN_CUDA(ctx, lg.alloc(4 * sK, s));
                CPN_CUDA(ctx, e1.alloc(sK, s)); CPN_CUDA(ctx, e2.alloc(sK, s)); CPN_CUDA(ctx, mml.alloc(sK, s)); CPN_CUDA(ctx, nme.alloc(sK, s));
                CPN_CUDA(ctx, cmd.alloc(sC, s)); CPN_CUDA(ctx, T.alloc(3 * sC, s));
                // the views: read-index rows, descriptors, RNG keys and the index tables of the summaries, built on the device
                { KTimer kt(ctx, CPN_K_PREP);
                  k_two_samp_views<<<dim3(blocks_for(std::max(Pmax, 1), 128), (unsigned)Rc), 128, 0, s>>>(
                      d_regs.p, sb.pk.rdesc.p, perm_ids ? d_pid.p : nullptr, ctx->seed, V, Pt, sN, sK, sC, sb.vdesc.p, sb.ridx.p, sb.vkey.p,
                      d_reg.p, d_exo.p, d_ko.p, d_co.p, d_pa.p, d_pb.p); }
                CPN_CUDA(ctx, cudaGetLastError());
                CPN_CUDA(ctx, cpn_wait_stream(ctx, s));               // the host staging vectors may go
                launches += 1;
            }
            // 2·ΣP multi-start EM fits over the views; results stay on the device
            rc = fit_impl(ctx, true, Rc, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, n_init, max_em_iters, amax, bmax,
                          cmax, mstep_mode, d_phi.p, d_Q.p, d_stat.p, nullptr, nullptr, nullptr, nullptr, &sb);
            if (rc) return rc;
            launches += ctx->last_launches;
            ctx->last_launches = 0;
            for (int k = 0; k < 4; ++k) work_acc[k] += ctx->last_work[k];
            // per-CpG summaries of every refit, CMD of the two halves of every permutation, then the three statistics
            {
                CpnArenaScope scope(&ctx->stage_arena);
                rc = cpn_summaries_and_cmd(ctx, ordered, V, d_phi.p, d_reg.p, d_exo.p, d_ko.p, sN, sK, Rc, sumN, d_cpg.p, d_rho.p, d_dn.p, d_so.p, d_sp.p,
                                           d_sq.p, Nmax, Pt, d_pa.p, d_pb.p, d_co.p, logZ.p, lg.p, e1.p, e2.p, mml.p, nme.p, cmd.p);
                if (rc) return rc;
            }
            { KTimer kt(ctx, CPN_K_SWEEP);
              k_two_samp_T<<<blocks_for(Pt * 32, 256), 256, 0, s>>>(Pt, d_ko.p, d_co.p, mml.p, nme.p, cmd.p, d_stat.p, T.p); }
            CPN_CUDA(ctx, cudaGetLastError());
            if (T_null) CPN_CUDA(ctx, T.download(T_null + 3 * tn[r0], 3 * sC, s));
            if (fit_status) CPN_CUDA(ctx, d_stat.download(fit_status + 2 * perm_off[r0], V, s));
            if (observed) {
                rc = two_samp_observed(ctx, Rc, phi_s1 + 3 * r0, phi_s2 + 3 * r0, po.data(), tn.data() + r0, exact ? exact + r0 : nullptr, cp.data(),
                                       so.data(), (int)Kmax, Nmax, d_cpg.p, d_rho.p, d_dn.p, d_so.p, d_sp.p, d_sq.p, T.p, T_obs + 3 * k0, cnt + 3 * k0,
                                       n_valid + 3 * k0, pval + 3 * k0, ordered);
                if (rc) return rc;
            }
            CPN_CUDA(ctx, cpn_wait_stream(ctx, s));
            launches += ctx->last_launches;
        } else if (observed) {
            // no permutations at all in this chunk (binomial ≤ 20 or pval_comp off): observed statistics only, p-values NaN
            CPN_CUDA(ctx, ctx->stage_arena.reset());
            CpnArenaScope scope(&ctx->stage_arena);
            DevBuf<int64_t> d_cpg, d_so, d_sp, d_sq;
            DevBuf<double> d_rho, d_dn;
            CPN_CUDA(ctx, d_cpg.upload(cp.data(), Rc + 1, s)); CPN_CUDA(ctx, d_so.upload(so.data(), Rc + 1, s));
            CPN_CUDA(ctx, d_sp.upload(sub_p + k0, sumK, s)); CPN_CUDA(ctx, d_sq.upload(sub_q + k0, sumK, s));
            CPN_CUDA(ctx, d_rho.upload(rho_n + n0, sumN, s)); CPN_CUDA(ctx, d_dn.upload(d_n + (n0 - r0), sumN - Rc, s));
            ctx->last_launches = 0;
            rc = two_samp_observed(ctx, Rc, phi_s1 + 3 * r0, phi_s2 + 3 * r0, po.data(), tn.data() + r0, exact ? exact + r0 : nullptr, cp.data(),
                                   so.data(), (int)Kmax, Nmax, d_cpg.p, d_rho.p, d_dn.p, d_so.p, d_sp.p, d_sq.p, nullptr, T_obs + 3 * k0, cnt + 3 * k0,
                                   n_valid + 3 * k0, pval + 3 * k0, ordered);
            if (rc) return rc;
            CPN_CUDA(ctx, cpn_wait_stream(ctx, s));
            launches += ctx->last_launches;
        }
    }
    CPN_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    CPN_CUDA(ctx, cpn_wait_stream(ctx, s));
    float ms = 0;
    CPN_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    ctx->last_launches = launches;
    for (int k = 0; k < 4; ++k) ctx->last_work[k] = work_acc[k];
    if (ctx->profile) cpn_profile_collect(ctx);
    return CPN_OK;
}

int cpn_two_samp_null_batch(cpn_ctx* ctx, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                            const int64_t* read_off, const double* log_pyx_u, const double* log_pyx_m, const int64_t* perm_off,
                            const int64_t* n1, const int32_t* perm_ids, const double* phi0, int32_t n_init, int32_t max_em_iters,
                            double amax, double bmax, double cmax, int32_t mstep_mode, const int64_t* cpg_off, const double* rho_n,
                            const double* d_n, const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q, double* T_null,
                            int32_t* fit_status) {
    if (!ctx) return CPN_ERR_ARG;
    CPN_ARG(ctx, T_null, "null pointer");
    return two_samp_impl(ctx, R, grp_off, Nl, rho_l, d_l, read_off, log_pyx_u, log_pyx_m, n1, nullptr, nullptr, perm_off, perm_ids, nullptr, phi0,
                         n_init, max_em_iters, amax, bmax, cmax, mstep_mode, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, nullptr, nullptr, nullptr,
                         nullptr, T_null, fit_status);
}

int cpn_two_samp_test_batch(cpn_ctx* ctx, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                            const int64_t* read_off, const double* log_pyx_u, const double* log_pyx_m, const int64_t* n1, const double* phi_s1,
                            const double* phi_s2, const int64_t* perm_off, const int32_t* perm_ids, const int32_t* exact, const double* phi0,
                            int32_t n_init, int32_t max_em_iters, double amax, double bmax, double cmax, int32_t mstep_mode,
                            const int64_t* cpg_off, const double* rho_n, const double* d_n, const int64_t* sub_off, const int64_t* sub_p,
                            const int64_t* sub_q, double* T_obs, int64_t* cnt, int64_t* n_valid, double* pval, double* T_null,
                            int32_t* fit_status) {
    if (!ctx) return CPN_ERR_ARG;
    CPN_ARG(ctx, phi_s1 && phi_s2, "null pointer");
    if (ctx->devs.size() <= 1 || R < 2)
        return two_samp_impl(ctx, R, grp_off, Nl, rho_l, d_l, read_off, log_pyx_u, log_pyx_m, n1, phi_s1, phi_s2, perm_off, perm_ids, exact, phi0,
                             n_init, max_em_iters, amax, bmax, cmax, mstep_mode, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, T_obs, cnt, n_valid, pval,
                             T_null, fit_status);
    // cpn_create_multi: contiguous shards of regions balanced by permutations x reads x groups, one host thread per device
    CPN_ARG(ctx, grp_off && read_off && perm_off && n1 && cpg_off && sub_off, "null pointer");
    const int nd = (int)ctx->devs.size();
    std::vector<int64_t> obs(R + 1), pid(R + 1), tn(R + 1);
    std::vector<double> cost(R + 1);
    obs[0] = pid[0] = tn[0] = 0; cost[0] = 0;
    for (int64_t r = 0; r < R; ++r) {
        const int64_t L = grp_off[r + 1] - grp_off[r], n = read_off[r + 1] - read_off[r], P = perm_off[r + 1] - perm_off[r];
        obs[r + 1] = obs[r] + L * n; pid[r + 1] = pid[r] + P * n1[r]; tn[r + 1] = tn[r] + P * (sub_off[r + 1] - sub_off[r]);
        cost[r + 1] = cost[r] + (double)(P + 1) * (double)(L * n);
    }
    std::vector<int64_t> bnd(nd + 1, R);
    bnd[0] = 0;
    for (int d = 1; d < nd; ++d) {
        bnd[d] = std::lower_bound(cost.begin(), cost.end(), cost[R] * d / nd) - cost.begin();
        bnd[d] = std::min(std::max(bnd[d], bnd[d - 1]), R);
    }
    std::vector<int> rc(nd, CPN_OK);
    std::vector<std::thread> th;
    for (int d = 0; d < nd; ++d) {
        th.emplace_back([&, d]() {
            cpn_ctx* c = ctx->devs[d];
            const int64_t r0 = bnd[d], Rd = bnd[d + 1] - r0;
            c->last_ms = 0; c->last_launches = 0;
            for (double& w : c->last_work) w = 0;
            if (Rd <= 0) return;
            c->seed = ctx->seed; c->region_base = ctx->region_base + r0;
            if (d > 0) cpn_set_profile(c, ctx->profile ? 1 : 0);
            std::vector<int64_t> g(Rd + 1), rd(Rd + 1), cp(Rd + 1), so(Rd + 1), po(Rd + 1);
            for (int64_t i = 0; i <= Rd; ++i) {
                g[i] = grp_off[r0 + i] - grp_off[r0]; rd[i] = read_off[r0 + i] - read_off[r0]; cp[i] = cpg_off[r0 + i] - cpg_off[r0];
                so[i] = sub_off[r0 + i] - sub_off[r0]; po[i] = perm_off[r0 + i] - perm_off[r0];
            }
            const int64_t g0 = grp_off[r0], n0 = cpg_off[r0], k0 = sub_off[r0], p0 = perm_off[r0];
            rc[d] = two_samp_impl(c, Rd, g.data(), Nl + g0, rho_l + g0, d_l + (g0 - r0), rd.data(), log_pyx_u + obs[r0], log_pyx_m + obs[r0], n1 + r0,
                                  phi_s1 + 3 * r0, phi_s2 + 3 * r0, po.data(), perm_ids ? perm_ids + pid[r0] : nullptr, exact ? exact + r0 : nullptr,
                                  phi0 ? phi0 + 2 * p0 * n_init * 3 : nullptr, n_init, max_em_iters, amax, bmax, cmax, mstep_mode, cp.data(), rho_n + n0,
                                  d_n + (n0 - r0), so.data(), sub_p + k0, sub_q + k0, T_obs + 3 * k0, cnt + 3 * k0, n_valid + 3 * k0, pval + 3 * k0,
                                  T_null ? T_null + 3 * tn[r0] : nullptr, fit_status ? fit_status + 2 * p0 : nullptr);
            c->region_base = 0;
        });
    }
    for (auto& t : th) t.join();
    cudaSetDevice(ctx->device);
    int ret = CPN_OK;
    double ms = 0, work[4] = {0, 0, 0, 0};
    int64_t launches = 0;
    for (int d = 0; d < nd; ++d) {
        cpn_ctx* c = ctx->devs[d];
        if (rc[d] != CPN_OK && ret == CPN_OK) { ret = rc[d]; if (d > 0) ctx->err = c->err; }
        ms = std::max(ms, c->last_ms); launches += c->last_launches;
        for (int k = 0; k < 4; ++k) work[k] += c->last_work[k];
        if (d > 0)
            for (int k = 0; k < CPN_K_NCLASS; ++k) {
                ctx->prof_ms[k] += c->prof_ms[k]; ctx->prof_launches[k] += c->prof_launches[k];
                c->prof_ms[k] = 0; c->prof_launches[k] = 0;
            }
    }
    ctx->last_ms = ms; ctx->last_launches = launches;
    for (int k = 0; k < 4; ++k) ctx->last_work[k] = work[k];
    return ret;
}

}  // extern "C"

