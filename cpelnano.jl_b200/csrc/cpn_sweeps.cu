// cpn_sweeps.cu — permutation-test statistic sweeps with exact integer counts.
//
// Reference path:
//   unmatched groups  unmat_est_reg_test            src/grp_perm_tests.jl:200-291 (stats :13-189)
//   matched groups    mat_est_reg_test              src/grp_perm_tests.jl:432-499 (stats :305-421)
//   two samples       p-value counting              src/two_samp_perm_tests.jl:230-258
//
// One block per estimation region: the region's statistic tables are staged in shared memory, threads
// own labelings (or sign patterns / permutations) and loop over the region's sub-regions; counts are
// integer shared-memory atomics, so they do not depend on thread scheduling.  Every floating-point
// operation of a statistic uses the explicitly rounded intrinsics (__dadd_rn, __ddiv_rn, …) in the
// reference's accumulation order, so nvcc cannot contract or re-associate them: given bit-identical
// tables the counts equal the reference's.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <thread>
#include <vector>

#include "cpn_common.cuh"

namespace {

constexpr int SW_THREADS = 256;
inline unsigned blocks_for(int64_t n, int per_block) { return (unsigned)((n + per_block - 1) / per_block); }

// a / n for a small positive integer n, correctly rounded, in three FP64 instructions instead of the ≈ 25 of the generic
// division: q0 = RN(a·y) with y = RN(1/n), r = a − n·q0 (exact in an FMA), q = RN(q0 + r·y) (Markstein's final correction;
// checked bit-for-bit against a/n on 4·10⁸ values for every n ≤ 1024).  Outside the range where neither step can
// underflow or overflow — and for NaN — the generic division runs.
__device__ __forceinline__ double div_small(double a, double n, double y) {
    const double aa = fabs(a);
    if (!(aa >= 1e-290 && aa <= 1e290) && a != 0.0) return __ddiv_rn(a, n);
    const double q0 = __dmul_rn(a, y);
    const double r = __fma_rn(-n, q0, a);
    return __fma_rn(r, y, q0);
}

// Unmatched sweep.  A thread owns labelings; for each it keeps the group membership as two bit masks and KB sub-regions'
// worth of running sums in registers, so the index arithmetic of a sample (or pair of samples) is paid once per KB
// sub-regions and the inner loops are LDS + DADD.  Sums run over ascending sample ids — the order of the reference's id
// vectors (grp_perm_tests.jl:121-125, :86-107; combinations are ascending, and so is their complement).
constexpr int SW_PK_MAX = 1024;   // labelings whose packed member lists fit the kernel's shared-memory table

// PACKED (n1 + n2 ≤ 16 samples and at most SW_PK_MAX labelings — every exact test and the reference's LMAX = 1000 sample): a
// labeling is held as one 64-bit word of 4-bit sample ids, group 1 ascending then group 2 ascending, built once per block; the
// sweep then walks members with a shift and a mask instead of a find-first-set chain over two bit masks, runs over sub-region
// passes in the OUTER loop so that a thread counts its exceedances in registers, and touches the shared counters once per warp
// and pass (a ballot-free shuffle sum) instead of once per (labeling, sub-region, statistic).  Same additions in the same order.
template <int KB, bool PACKED>
__global__ void __launch_bounds__(SW_THREADS) k_unmat_sweep(int n1, int n2, int Kmax, const int64_t* __restrict__ tbl_off,
                                                             const double* __restrict__ mml, const double* __restrict__ nme,
                                                             const double* __restrict__ cmd_tbl, const int32_t* __restrict__ labelings,
                                                             int64_t P, int exact, double* __restrict__ T, int64_t* __restrict__ cnt,
                                                             double* __restrict__ pval) {
    extern __shared__ double sm[];
    const int S = n1 + n2;
    int64_t r = blockIdx.x;
    int64_t k0 = tbl_off[r];
    int K = (int)(tbl_off[r + 1] - k0);
    double* s_mml = sm;                          // [S][K]
    double* s_nme = s_mml + S * Kmax;            // [S][K]
    double* s_cmd = s_nme + S * Kmax;            // [S*S][K]
    double* s_T = s_cmd + S * S * Kmax;          // [3][K]
    int* s_cnt = (int*)(s_T + 3 * Kmax);         // [3][K]  (+ 16 doubles of slack: row tails are read unguarded below)
    unsigned long long* s_pk = (unsigned long long*)(sm + (2 * S + S * S + 3) * Kmax + 16 + (3 * Kmax + 1) / 2);   // [P] (PACKED)
    for (int i = threadIdx.x; i < S * K; i += blockDim.x) { s_mml[i] = mml[k0 * S + i]; s_nme[i] = nme[k0 * S + i]; }
    if (PACKED) {
        for (int64_t pi = threadIdx.x; pi < P; pi += blockDim.x) {
            const int32_t* lab = labelings + pi * n1;
            unsigned m1 = 0;
            unsigned long long pk = 0;
            for (int i = 0; i < n1; ++i) m1 |= 1u << (lab[i] - 1);
            int pos1 = 0, pos2 = n1;                     // both groups in ascending id order, whatever the order of the row
            for (int id = 0; id < S; ++id) {
                if ((m1 >> id) & 1u) { pk |= (unsigned long long)id << (4 * pos1); ++pos1; }
                else { pk |= (unsigned long long)id << (4 * pos2); ++pos2; }
            }
            s_pk[pi] = pk;
        }
    }
    for (int i = threadIdx.x; i < S * S * K; i += blockDim.x) s_cmd[i] = cmd_tbl[k0 * S * S + i];
    for (int i = threadIdx.x; i < 3 * K; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    // observed statistics (grp_perm_tests.jl:212-214)
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double a1 = s_mml[k], b1 = s_nme[k];
        for (int s = 1; s < n1; ++s) { a1 = __dadd_rn(a1, s_mml[s * K + k]); b1 = __dadd_rn(b1, s_nme[s * K + k]); }
        double a2 = s_mml[n1 * K + k], b2 = s_nme[n1 * K + k];
        for (int s = n1 + 1; s < S; ++s) { a2 = __dadd_rn(a2, s_mml[s * K + k]); b2 = __dadd_rn(b2, s_nme[s * K + k]); }
        s_T[k] = __dsub_rn(__ddiv_rn(a1, (double)n1), __ddiv_rn(a2, (double)n2));
        s_T[K + k] = __dsub_rn(__ddiv_rn(b1, (double)n1), __ddiv_rn(b2, (double)n2));
        double acc = s_cmd[(0 * S + n1) * K + k];
        for (int i = 0; i < n1; ++i)
            for (int j = 0; j < n2; ++j)
                if (i | j) acc = __dadd_rn(acc, s_cmd[(i * S + n1 + j) * K + k]);
        s_T[2 * K + k] = __ddiv_rn(acc, (double)(n1 * n2));
    }
    __syncthreads();
    const unsigned long long all = (S >= 64) ? ~0ull : ((1ull << S) - 1ull);
    const double d1 = (double)n1, d2 = (double)n2, d12 = (double)(n1 * n2);
    const double y1 = __ddiv_rn(1.0, d1), y2 = __ddiv_rn(1.0, d2), y12 = __ddiv_rn(1.0, d12);
    if (PACKED) {
        const int lane = threadIdx.x & 31;
        for (int kb = 0; kb < K; kb += KB) {
            int cm[KB], cn[KB], cc[KB];
#pragma unroll
            for (int u = 0; u < KB; ++u) { cm[u] = 0; cn[u] = 0; cc[u] = 0; }
            for (int64_t pi = threadIdx.x; pi < P; pi += blockDim.x) {
                const unsigned long long pk = s_pk[pi];
                double tm[KB], tn[KB], tc[KB];
                {   // Tmml, Tnme: mean over group 1 − mean over group 2, members in ascending id order
                    double a[KB], b[KB];
                    int o = (int)(pk & 15ull) * K + kb;
#pragma unroll
                    for (int u = 0; u < KB; ++u) { a[u] = s_mml[o + u]; b[u] = s_nme[o + u]; }
#pragma unroll 4
                    for (int i = 1; i < n1; ++i) {
                        o = (int)((pk >> (4 * i)) & 15ull) * K + kb;
#pragma unroll
                        for (int u = 0; u < KB; ++u) { a[u] = __dadd_rn(a[u], s_mml[o + u]); b[u] = __dadd_rn(b[u], s_nme[o + u]); }
                    }
#pragma unroll
                    for (int u = 0; u < KB; ++u) { tm[u] = div_small(a[u], d1, y1); tn[u] = div_small(b[u], d1, y1); }
                    o = (int)((pk >> (4 * n1)) & 15ull) * K + kb;
#pragma unroll
                    for (int u = 0; u < KB; ++u) { a[u] = s_mml[o + u]; b[u] = s_nme[o + u]; }
#pragma unroll 4
                    for (int i = n1 + 1; i < S; ++i) {
                        o = (int)((pk >> (4 * i)) & 15ull) * K + kb;
#pragma unroll
                        for (int u = 0; u < KB; ++u) { a[u] = __dadd_rn(a[u], s_mml[o + u]); b[u] = __dadd_rn(b[u], s_nme[o + u]); }
                    }
#pragma unroll
                    for (int u = 0; u < KB; ++u) {
                        tm[u] = __dsub_rn(tm[u], div_small(a[u], d2, y2));
                        tn[u] = __dsub_rn(tn[u], div_small(b[u], d2, y2));
                    }
                }
#pragma unroll
                for (int u = 0; u < KB; ++u) tc[u] = 0.0;                        // comp_unmat_perm_stat_cmd :86-107
                for (int i = 0; i < n1; ++i) {
                    const int ii = (int)((pk >> (4 * i)) & 15ull);
#pragma unroll 6
                    for (int j = n1; j < S; ++j) {
                        const int jj = (int)((pk >> (4 * j)) & 15ull);
                        const int o = (min(ii, jj) * S + max(ii, jj)) * K + kb;
#pragma unroll
                        for (int u = 0; u < KB; ++u) tc[u] = __dadd_rn(tc[u], s_cmd[o + u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < KB; ++u) {
                    const int k = kb + u;
                    if (k < K) {
                        const double tobs = s_T[k];
                        if (tobs == tobs) {                                      // isnan(tmml_obs[k]) && continue (:260)
                            cm[u] += fabs(tm[u]) >= fabs(tobs);                  // :268-270
                            cn[u] += fabs(tn[u]) >= fabs(s_T[K + k]);
                            cc[u] += div_small(tc[u], d12, y12) >= s_T[2 * K + k];
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < KB; ++u) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    cm[u] += __shfl_xor_sync(0xffffffffu, cm[u], o);
                    cn[u] += __shfl_xor_sync(0xffffffffu, cn[u], o);
                    cc[u] += __shfl_xor_sync(0xffffffffu, cc[u], o);
                }
                const int k = kb + u;
                if (lane == 0 && k < K) { atomicAdd(&s_cnt[k], cm[u]); atomicAdd(&s_cnt[K + k], cn[u]); atomicAdd(&s_cnt[2 * K + k], cc[u]); }
            }
        }
    } else
    for (int64_t pi = threadIdx.x; pi < P; pi += blockDim.x) {
        const int32_t* lab = labelings + pi * n1;
        unsigned long long m1 = 0;
        for (int i = 0; i < n1; ++i) m1 |= 1ull << (lab[i] - 1);
        const unsigned long long m2 = all & ~m1;
        for (int kb = 0; kb < K; kb += KB) {
            double tm[KB], tn[KB], tc[KB];
            {   // Tmml, Tnme: mean over group 1 − mean over group 2
                double a[KB], b[KB];
                unsigned long long m = m1;
                const double *pm = s_mml + (__ffsll((long long)m) - 1) * K + kb, *pn = s_nme + (__ffsll((long long)m) - 1) * K + kb;
#pragma unroll
                for (int u = 0; u < KB; ++u) { a[u] = pm[u]; b[u] = pn[u]; }
                for (m &= m - 1; m; m &= m - 1) {
                    const int o = (__ffsll((long long)m) - 1) * K + kb;
#pragma unroll
                    for (int u = 0; u < KB; ++u) { a[u] = __dadd_rn(a[u], s_mml[o + u]); b[u] = __dadd_rn(b[u], s_nme[o + u]); }
                }
#pragma unroll
                for (int u = 0; u < KB; ++u) { tm[u] = div_small(a[u], d1, y1); tn[u] = div_small(b[u], d1, y1); }
                m = m2;
                pm = s_mml + (__ffsll((long long)m) - 1) * K + kb; pn = s_nme + (__ffsll((long long)m) - 1) * K + kb;
#pragma unroll
                for (int u = 0; u < KB; ++u) { a[u] = pm[u]; b[u] = pn[u]; }
                for (m &= m - 1; m; m &= m - 1) {
                    const int o = (__ffsll((long long)m) - 1) * K + kb;
#pragma unroll
                    for (int u = 0; u < KB; ++u) { a[u] = __dadd_rn(a[u], s_mml[o + u]); b[u] = __dadd_rn(b[u], s_nme[o + u]); }
                }
#pragma unroll
                for (int u = 0; u < KB; ++u) {
                    tm[u] = __dsub_rn(tm[u], div_small(a[u], d2, y2));
                    tn[u] = __dsub_rn(tn[u], div_small(b[u], d2, y2));
                }
            }
#pragma unroll
            for (int u = 0; u < KB; ++u) tc[u] = 0.0;                            // comp_unmat_perm_stat_cmd :86-107
            for (unsigned long long a = m1; a; a &= a - 1) {
                const int ii = __ffsll((long long)a) - 1;
                for (unsigned long long b = m2; b; b &= b - 1) {
                    const int jj = __ffsll((long long)b) - 1;
                    const int o = (min(ii, jj) * S + max(ii, jj)) * K + kb;
#pragma unroll
                    for (int u = 0; u < KB; ++u) tc[u] = __dadd_rn(tc[u], s_cmd[o + u]);
                }
            }
#pragma unroll
            for (int u = 0; u < KB; ++u) {
                const int k = kb + u;
                if (k < K) {
                    const double tobs = s_T[k];
                    if (tobs == tobs) {                                          // isnan(tmml_obs[k]) && continue (:260)
                        if (fabs(tm[u]) >= fabs(tobs)) atomicAdd(&s_cnt[k], 1);  // :268-270
                        if (fabs(tn[u]) >= fabs(s_T[K + k])) atomicAdd(&s_cnt[K + k], 1);
                        if (div_small(tc[u], d12, y12) >= s_T[2 * K + k]) atomicAdd(&s_cnt[2 * K + k], 1);
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * K; i += blockDim.x) {
        int t = i / K, k = i % K;
        int64_t o = 3 * k0 + (int64_t)t * K + k;
        double tobs = s_T[k];
        T[o] = s_T[i];
        cnt[o] = s_cnt[i];
        double p = cpn_nan();
        if (P > 0 && tobs == tobs)
            p = exact ? __ddiv_rn((double)s_cnt[i], (double)P) : __ddiv_rn(__dadd_rn(1.0, (double)s_cnt[i]), __dadd_rn(1.0, (double)P));
        pval[o] = p;
    }
}

typedef void (*unmat_kernel_t)(int, int, int, const int64_t*, const double*, const double*, const double*, const int32_t*, int64_t, int, double*,
                               int64_t*, double*);
// Sub-regions per pass: the fewest passes of at most 8, split evenly (K = 9 → 5 + 4), so the accumulators stay in ≈ 100
// registers and two blocks fit on an SM.  CPN_SWEEP_KB overrides (4, 5, 6, 8, 12, 16) for experiments.
inline bool unmat_packed(int S, int64_t P) {
    const char* e = getenv("CPN_SWEEP_PACKED");
    return S <= 16 && P <= SW_PK_MAX && !(e && atoi(e) == 0);
}
inline unmat_kernel_t unmat_kernel_for(int Kmax, int S, int64_t P) {
    int chunks = (Kmax + 7) / 8, kb = (Kmax + chunks - 1) / chunks;
    if (const char* e = getenv("CPN_SWEEP_KB")) kb = atoi(e);
    if (unmat_packed(S, P))
        return kb <= 4 ? k_unmat_sweep<4, true> : kb == 5 ? k_unmat_sweep<5, true> : kb == 6 ? k_unmat_sweep<6, true>
             : kb <= 8 ? k_unmat_sweep<8, true> : kb <= 12 ? k_unmat_sweep<12, true> : k_unmat_sweep<16, true>;
    return kb <= 4 ? k_unmat_sweep<4, false> : kb == 5 ? k_unmat_sweep<5, false> : kb == 6 ? k_unmat_sweep<6, false>
         : kb <= 8 ? k_unmat_sweep<8, false> : kb <= 12 ? k_unmat_sweep<12, false> : k_unmat_sweep<16, false>;
}
inline size_t unmat_smem(int S, int Kmax, int64_t P) {
    size_t doubles = (size_t)(2 * S + S * S + 3) * Kmax + 16 + (size_t)(3 * Kmax + 1) / 2;
    if (unmat_packed(S, P)) doubles += (size_t)P;
    return doubles * sizeof(double);
}

__global__ void __launch_bounds__(SW_THREADS) k_mat_sweep(int n, int Kmax, const int64_t* __restrict__ tbl_off, const double* __restrict__ mml1,
                                                           const double* __restrict__ mml2, const double* __restrict__ nme1,
                                                           const double* __restrict__ nme2, const int64_t* __restrict__ js, int64_t P,
                                                           int exact, double* __restrict__ T, int64_t* __restrict__ cnt,
                                                           double* __restrict__ pval, int rstride /*values per sub-region and region: n, or S
                                                           when both groups share one [R][S][K] table*/, int g2off /*0, or n*/) {
    extern __shared__ double sm[];
    int64_t r = blockIdx.x;
    int64_t k0 = tbl_off[r];
    int K = (int)(tbl_off[r + 1] - k0);
    double* s_dm = sm;                          // [n][K] pair differences (comp_mat_diff_mml :343)
    double* s_dn = s_dm + n * Kmax;
    double* s_T = s_dn + n * Kmax;              // [2][K]
    int* s_cnt = (int*)(s_T + 2 * Kmax);
    for (int i = threadIdx.x; i < n * K; i += blockDim.x) {
        s_dm[i] = __dsub_rn(mml1[k0 * rstride + i], mml2[k0 * rstride + g2off * K + i]);
        s_dn[i] = __dsub_rn(nme1[k0 * rstride + i], nme2[k0 * rstride + g2off * K + i]);
    }
    for (int i = threadIdx.x; i < 2 * K; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double a = s_dm[k], b = s_dn[k];
        for (int s = 1; s < n; ++s) { a = __dadd_rn(a, s_dm[s * K + k]); b = __dadd_rn(b, s_dn[s * K + k]); }
        s_T[k] = __ddiv_rn(a, (double)n);                                        // :448-449
        s_T[K + k] = __ddiv_rn(b, (double)n);
    }
    __syncthreads();
    const double dn_ = (double)n, yn = __ddiv_rn(1.0, (double)n);
    for (int64_t pi = threadIdx.x; pi < P; pi += blockDim.x) {
        int64_t j = js[pi];
        for (int k = 0; k < K; ++k) {
            double tobs = s_T[k];
            if (tobs != tobs) continue;
            int64_t jk = j;
            double a = 0.0, b = 0.0;                                             // comp_mat_j_stat :388-405
            for (int s = 0; s < n; ++s) {
                bool bit = (jk & 1) != 0;
                a = __dadd_rn(a, bit ? s_dm[s * K + k] : -s_dm[s * K + k]);
                b = __dadd_rn(b, bit ? s_dn[s * K + k] : -s_dn[s * K + k]);
                jk >>= 1;
            }
            a = div_small(a, dn_, yn); b = div_small(b, dn_, yn);
            if (fabs(a) >= fabs(tobs)) atomicAdd(&s_cnt[k], 1);
            if (fabs(b) >= fabs(s_T[K + k])) atomicAdd(&s_cnt[K + k], 1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * K; i += blockDim.x) {
        int t = i / K, k = i % K;
        int64_t o = 2 * k0 + (int64_t)t * K + k;
        double tobs = s_T[k];
        T[o] = s_T[i];
        cnt[o] = s_cnt[i];
        double p = cpn_nan();
        if (P > 0 && tobs == tobs)
            p = exact ? __ddiv_rn((double)s_cnt[i], (double)P) : __ddiv_rn(__dadd_rn(1.0, (double)s_cnt[i]), __dadd_rn(1.0, (double)P));
        pval[o] = p;
    }
}

// P permutations for every region (perm_off == null), or region r's own count perm_off[r+1] − perm_off[r] with its null
// statistics at T_null + 3·nul_off[r]; exact_r (optional) overrides `exact` per region.
__global__ void __launch_bounds__(SW_THREADS) k_two_samp_pvals(int Kmax, const int64_t* __restrict__ tbl_off, int64_t P,
                                                                const double* __restrict__ T_obs, const double* __restrict__ T_null, int exact,
                                                                int64_t* __restrict__ cnt, int64_t* __restrict__ n_valid, double* __restrict__ pval,
                                                                const int64_t* __restrict__ perm_off, const int64_t* __restrict__ nul_off,
                                                                const int32_t* __restrict__ exact_r) {
    extern __shared__ double sm[];
    int64_t r = blockIdx.x;
    int64_t k0 = tbl_off[r];
    int K = (int)(tbl_off[r + 1] - k0);
    double* s_T = sm;                            // [3][K]
    int* s_cnt = (int*)(s_T + 3 * Kmax);         // [3][K]
    int* s_nv = s_cnt + 3 * Kmax;                // [3][K]
    for (int i = threadIdx.x; i < 3 * K; i += blockDim.x) { s_T[i] = T_obs[3 * k0 + i]; s_cnt[i] = 0; s_nv[i] = 0; }
    __syncthreads();
    const double* nul = T_null + (perm_off ? 3 * nul_off[r] : 3 * k0 * P);
    if (perm_off) P = perm_off[r + 1] - perm_off[r];
    if (exact_r) exact = exact_r[r];
    for (int64_t i = threadIdx.x; i < P * 3 * K; i += blockDim.x) {
        int64_t rem = i % (3 * K);
        int t = (int)(rem / K), k = (int)(rem % K);
        if (s_T[k] != s_T[k]) continue;                                          // isnan(tmml_obs[k]) && continue (:233)
        double v = nul[i];
        if (v != v) continue;                                                    // drop NaN nulls (:241-243)
        atomicAdd(&s_nv[t * K + k], 1);
        bool ge = (t < 2) ? (fabs(v) >= fabs(s_T[t * K + k])) : (v >= s_T[t * K + k]);
        if (ge) atomicAdd(&s_cnt[t * K + k], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * K; i += blockDim.x) {
        int k = i % K;
        int64_t o = 3 * k0 + i;
        cnt[o] = s_cnt[i];
        n_valid[o] = s_nv[i];
        double p = cpn_nan();
        if (s_T[k] == s_T[k] && s_nv[k] > 20)                                    // length(tmml_perms_k) > 20 || continue (:246)
            p = exact ? __ddiv_rn((double)s_cnt[i], (double)s_nv[i]) : __ddiv_rn(__dadd_rn(1.0, (double)s_cnt[i]), __dadd_rn(1.0, (double)s_nv[i]));
        pval[o] = p;
    }
}

// Fused group test: index tables of the R·S models (model t = r·S + s shares region r's features).
__global__ void k_grp_models(int64_t T, int S, const int64_t* __restrict__ cpg_off, const int64_t* __restrict__ sub_off,
                             int64_t* __restrict__ reg_of, int64_t* __restrict__ ex_off, int64_t* __restrict__ k_off) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t > T) return;
    if (t == T) { int64_t R = T / S; ex_off[T] = S * cpg_off[R]; k_off[T] = S * sub_off[R]; return; }
    int64_t r = t / S, sidx = t - r * S;
    reg_of[t] = r;
    ex_off[t] = S * cpg_off[r] + sidx * (cpg_off[r + 1] - cpg_off[r]);
    k_off[t] = S * sub_off[r] + sidx * (sub_off[r + 1] - sub_off[r]);
}
// Pairs of a region and where their CMD vectors go.  Unmatched: all i<j, written straight into the [S][S][K] table the
// sweep reads; matched: (s, n+s), written as [n][K].
__global__ void k_grp_pairs(int64_t R, int S, int npair, const int* __restrict__ pi, const int* __restrict__ pj, int matched,
                            const int64_t* __restrict__ sub_off, int64_t* __restrict__ pa, int64_t* __restrict__ pb,
                            int64_t* __restrict__ cmd_off) {
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= R * npair) return;
    int64_t r = q / npair;
    int loc = (int)(q - r * npair);
    int i = pi[loc], j = pj[loc];
    int64_t k0 = sub_off[r], K = sub_off[r + 1] - k0;
    pa[q] = r * S + i; pb[q] = r * S + j;
    cmd_off[q] = matched ? (int64_t)npair * k0 + (int64_t)loc * K : (int64_t)S * S * k0 + (int64_t)(i * S + j) * K;
}
// Tcmd of the matched test: sum(cmds)/n, left to right (grp_perm_tests.jl:308-314).
__global__ void k_mat_tcmd(int64_t R, int n, const int64_t* __restrict__ sub_off, const double* __restrict__ cmd /*[R][n][K]*/,
                           double* __restrict__ tcmd) {
    int64_t r = blockIdx.x;
    int64_t k0 = sub_off[r];
    int K = (int)(sub_off[r + 1] - k0);
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const double* c = cmd + (int64_t)n * k0;
        double a = c[k];
        for (int s = 1; s < n; ++s) a = __dadd_rn(a, c[(int64_t)s * K + k]);
        tcmd[k0 + k] = __ddiv_rn(a, (double)n);
    }
}

int kmax_of(const int64_t* tbl_off, int64_t R) {
    int64_t k = 1;
    for (int64_t r = 0; r < R; ++r) k = std::max(k, tbl_off[r + 1] - tbl_off[r]);
    return (int)k;
}

}  // namespace

// Launch-only entry of the two-sample counting kernel (all pointers on the device), for the fused two-sample test.
int cpn_two_samp_pvals_launch(cpn_ctx* ctx, int64_t R, int Kmax, const int64_t* tbl_off, const int64_t* perm_off, const int64_t* nul_off,
                              const int32_t* exact_r, const double* T_obs, const double* T_null, int64_t* cnt, int64_t* n_valid, double* pval) {
    size_t smem = (size_t)3 * Kmax * sizeof(double) + (size_t)6 * Kmax * sizeof(int);
    {
        KTimer kt(ctx, CPN_K_SWEEP);
        k_two_samp_pvals<<<(unsigned)R, SW_THREADS, smem, ctx->stream>>>(Kmax, tbl_off, 0, T_obs, T_null, 0, cnt, n_valid, pval, perm_off, nul_off, exact_r);
    }
    CPN_CUDA(ctx, cudaGetLastError());
    return CPN_OK;
}

extern "C" {

int cpn_grp_unmat_sweep(cpn_ctx* ctx, int64_t R, int32_t n1, int32_t n2, const int64_t* tbl_off, const double* mml, const double* nme,
                        const double* cmd_tbl, const int32_t* labelings, int64_t P, int32_t exact, double* T, int64_t* cnt, double* pval) {
    if (!ctx) return CPN_ERR_ARG;
    CPN_ARG(ctx, R >= 0 && n1 >= 1 && n2 >= 1 && n1 <= 32 && n2 <= 32 && P >= 0, "sizes (n1,n2 in 1..32)");
    CPN_ARG(ctx, tbl_off && mml && nme && cmd_tbl && (labelings || P == 0) && T && cnt && pval, "null pointer");
    CPN_CUDA(ctx, cudaSetDevice(ctx->device));
    LaunchCounter lc(ctx);
    if (R == 0) return CPN_OK;
    const int S = n1 + n2;
    const int Kmax = kmax_of(tbl_off, R);
    const int64_t sumK = tbl_off[R];
    size_t smem = unmat_smem(S, Kmax, P);
    CPN_ARG(ctx, smem <= 200 * 1024, "statistic tables of one region do not fit in shared memory");
    for (int64_t i = 0; i < P * n1; ++i) CPN_ARG(ctx, labelings[i] >= 1 && labelings[i] <= S, "labeling id out of range");
    cudaStream_t s = ctx->stream;
    unmat_kernel_t kern = unmat_kernel_for(Kmax, S, P);
    CPN_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CPN_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    {
        DevBuf<int64_t> d_off, o_cnt;
        DevBuf<double> d_mml, d_nme, d_cmd, o_T, o_p;
        DevBuf<int32_t> d_lab;
        CPN_CUDA(ctx, d_off.upload(tbl_off, R + 1, s)); CPN_CUDA(ctx, d_mml.upload(mml, sumK * S, s)); CPN_CUDA(ctx, d_nme.upload(nme, sumK * S, s));
        CPN_CUDA(ctx, d_cmd.upload(cmd_tbl, sumK * S * S, s)); CPN_CUDA(ctx, d_lab.upload(labelings, P * n1, s));
        CPN_CUDA(ctx, o_T.alloc(3 * sumK, s)); CPN_CUDA(ctx, o_cnt.alloc(3 * sumK, s)); CPN_CUDA(ctx, o_p.alloc(3 * sumK, s));
        { KTimer kt(ctx, CPN_K_SWEEP); kern<<<(unsigned)R, SW_THREADS, smem, s>>>(n1, n2, Kmax, d_off.p, d_mml.p, d_nme.p, d_cmd.p, d_lab.p, P, exact, o_T.p, o_cnt.p, o_p.p); }
        CPN_CUDA(ctx, cudaGetLastError());
        CPN_CUDA(ctx, o_T.download(T, 3 * sumK, s)); CPN_CUDA(ctx, o_cnt.download(cnt, 3 * sumK, s)); CPN_CUDA(ctx, o_p.download(pval, 3 * sumK, s));
        CPN_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
        CPN_CUDA(ctx, cudaStreamSynchronize(s));
    }
    float ms = 0;
    CPN_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    if (ctx->profile) cpn_profile_collect(ctx);
    return CPN_OK;
}

int cpn_grp_mat_sweep(cpn_ctx* ctx, int64_t R, int32_t n, const int64_t* tbl_off, const double* mml1, const double* mml2, const double* nme1,
                      const double* nme2, const int64_t* js, int64_t P, int32_t exact, double* T, int64_t* cnt, double* pval) {
    if (!ctx) return CPN_ERR_ARG;
    CPN_ARG(ctx, R >= 0 && n >= 1 && n <= 62 && P >= 0, "sizes (n in 1..62)");
    CPN_ARG(ctx, tbl_off && mml1 && mml2 && nme1 && nme2 && (js || P == 0) && T && cnt && pval, "null pointer");
    CPN_CUDA(ctx, cudaSetDevice(ctx->device));
    LaunchCounter lc(ctx);
    if (R == 0) return CPN_OK;
    const int Kmax = kmax_of(tbl_off, R);
    const int64_t sumK = tbl_off[R];
    size_t smem = (size_t)(2 * n + 2) * Kmax * sizeof(double) + (size_t)2 * Kmax * sizeof(int);
    CPN_ARG(ctx, smem <= 200 * 1024, "statistic tables of one region do not fit in shared memory");
    cudaStream_t s = ctx->stream;
    CPN_CUDA(ctx, cudaFuncSetAttribute(k_mat_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CPN_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    {
        DevBuf<int64_t> d_off, d_js, o_cnt;
        DevBuf<double> d_m1, d_m2, d_n1, d_n2, o_T, o_p;
        CPN_CUDA(ctx, d_off.upload(tbl_off, R + 1, s)); CPN_CUDA(ctx, d_js.upload(js, P, s));
        CPN_CUDA(ctx, d_m1.upload(mml1, sumK * n, s)); CPN_CUDA(ctx, d_m2.upload(mml2, sumK * n, s));
        CPN_CUDA(ctx, d_n1.upload(nme1, sumK * n, s)); CPN_CUDA(ctx, d_n2.upload(nme2, sumK * n, s));
        CPN_CUDA(ctx, o_T.alloc(2 * sumK, s)); CPN_CUDA(ctx, o_cnt.alloc(2 * sumK, s)); CPN_CUDA(ctx, o_p.alloc(2 * sumK, s));
        { KTimer kt(ctx, CPN_K_SWEEP); k_mat_sweep<<<(unsigned)R, SW_THREADS, smem, s>>>(n, Kmax, d_off.p, d_m1.p, d_m2.p, d_n1.p, d_n2.p, d_js.p, P, exact, o_T.p, o_cnt.p, o_p.p, n, 0); }
        CPN_CUDA(ctx, cudaGetLastError());
        CPN_CUDA(ctx, o_T.download(T, 2 * sumK, s)); CPN_CUDA(ctx, o_cnt.download(cnt, 2 * sumK, s)); CPN_CUDA(ctx, o_p.download(pval, 2 * sumK, s));
        CPN_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
        CPN_CUDA(ctx, cudaStreamSynchronize(s));
    }
    float ms = 0;
    CPN_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    if (ctx->profile) cpn_profile_collect(ctx);
    return CPN_OK;
}

int cpn_two_samp_pvals(cpn_ctx* ctx, int64_t R, const int64_t* tbl_off, int64_t P, const double* T_obs, const double* T_null, int32_t exact,
                       int64_t* cnt, int64_t* n_valid, double* pval) {
    if (!ctx) return CPN_ERR_ARG;
    CPN_ARG(ctx, R >= 0 && P >= 0 && tbl_off && T_obs && (T_null || P == 0) && cnt && n_valid && pval, "null pointer / sizes");
    CPN_CUDA(ctx, cudaSetDevice(ctx->device));
    LaunchCounter lc(ctx);
    if (R == 0) return CPN_OK;
    const int Kmax = kmax_of(tbl_off, R);
    const int64_t sumK = tbl_off[R];
    size_t smem = (size_t)3 * Kmax * sizeof(double) + (size_t)6 * Kmax * sizeof(int);
    cudaStream_t s = ctx->stream;
    CPN_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    {
        DevBuf<int64_t> d_off, o_cnt, o_nv;
        DevBuf<double> d_T, d_null, o_p;
        CPN_CUDA(ctx, d_off.upload(tbl_off, R + 1, s)); CPN_CUDA(ctx, d_T.upload(T_obs, 3 * sumK, s));
        CPN_CUDA(ctx, d_null.upload(T_null, 3 * sumK * P, s));
        CPN_CUDA(ctx, o_cnt.alloc(3 * sumK, s)); CPN_CUDA(ctx, o_nv.alloc(3 * sumK, s)); CPN_CUDA(ctx, o_p.alloc(3 * sumK, s));
        { KTimer kt(ctx, CPN_K_SWEEP); k_two_samp_pvals<<<(unsigned)R, SW_THREADS, smem, s>>>(Kmax, d_off.p, P, d_T.p, d_null.p, exact, o_cnt.p, o_nv.p, o_p.p, nullptr, nullptr, nullptr); }
        CPN_CUDA(ctx, cudaGetLastError());
        CPN_CUDA(ctx, o_cnt.download(cnt, 3 * sumK, s)); CPN_CUDA(ctx, o_nv.download(n_valid, 3 * sumK, s)); CPN_CUDA(ctx, o_p.download(pval, 3 * sumK, s));
        CPN_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
        CPN_CUDA(ctx, cudaStreamSynchronize(s));
    }
    float ms = 0;
    CPN_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    if (ctx->profile) cpn_profile_collect(ctx);
    return CPN_OK;
}

// Group comparison of R regions, in chunks of regions so that the workspace (≈ 25 KB per region at 6 vs 6, 9 sub-regions)
// does not grow with the batch: a whole genome (3·10⁶ regions) goes through the same few GB.  on_device: phi, the per-CpG tables
// and the outputs are device pointers; the CSR offsets and the labelings are host arrays either way (they size the launches).
static int grp_test_impl(cpn_ctx* ctx, bool on_device, int64_t R, int32_t n1, int32_t n2, int32_t matched, const double* phi,
                         const int64_t* cpg_off, const double* rho_n, const double* d_n, const int64_t* sub_off, const int64_t* sub_p,
                         const int64_t* sub_q, const void* labelings, int64_t P, int32_t exact, double* T, int64_t* cnt, double* pval,
                         double* tcmd) {
    CPN_ARG(ctx, R >= 0 && n1 >= 1 && n2 >= 1 && n1 <= 32 && n2 <= 32 && P >= 0, "sizes (n1,n2 in 1..32)");
    CPN_ARG(ctx, !matched || n1 == n2, "the matched test needs n1 == n2");
    CPN_ARG(ctx, phi && cpg_off && rho_n && d_n && sub_off && sub_p && sub_q && (labelings || P == 0) && T && cnt && pval, "null pointer");
    CPN_ARG(ctx, !matched || tcmd, "tcmd is required for the matched test");
    CPN_CUDA(ctx, cudaSetDevice(ctx->device));
    LaunchCounter lc(ctx);
    if (R == 0) return CPN_OK;
    const int S = n1 + n2, n = n1;
    const int Kmax = kmax_of(sub_off, R);
    int64_t Nmax = 0;
    for (int64_t r = 0; r < R; ++r) {
        CPN_ARG(ctx, cpg_off[r + 1] - cpg_off[r] >= 2, "every region needs N >= 2 CpG sites");
        Nmax = std::max(Nmax, cpg_off[r + 1] - cpg_off[r]);
    }
    std::vector<int> pi, pj;
    if (matched) for (int s = 0; s < n; ++s) { pi.push_back(s); pj.push_back(n + s); }
    else for (int i = 0; i < S; ++i) for (int j = i + 1; j < S; ++j) { pi.push_back(i); pj.push_back(j); }
    const int npair = (int)pi.size();
    const int nstat = matched ? 2 : 3;
    size_t smem = matched ? (size_t)(2 * n + 2) * Kmax * sizeof(double) + (size_t)2 * Kmax * sizeof(int) : unmat_smem(S, Kmax, P);
    CPN_ARG(ctx, smem <= 200 * 1024, "statistic tables of one region do not fit in shared memory");
    if (!matched) {
        const int32_t* lab = (const int32_t*)labelings;
        for (int64_t i = 0; i < P * n1; ++i) CPN_ARG(ctx, lab[i] >= 1 && lab[i] <= S, "labeling id out of range");
    }
    cudaStream_t s = ctx->stream;
    if (matched) CPN_CUDA(ctx, cudaFuncSetAttribute(k_mat_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unmat_kernel_t kern = unmat_kernel_for(Kmax, S, P);
    if (!matched) CPN_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const char* e_chunk = getenv("CPN_GRP_CHUNK_REGIONS");
    const int64_t chunk = std::max<int64_t>(1, e_chunk ? atoll(e_chunk) : 65536);
    const char* e_lane = getenv("CPN_CMD_LANE");
    const bool lane_ok = !(e_lane && atoi(e_lane) == 0);
    CPN_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    int64_t launches = 0;
    {
        // labelings and the pair list are shared by every chunk
        DevBuf<int64_t> d_js;
        DevBuf<int32_t> d_lab;
        DevBuf<int> d_pi, d_pj;
        CPN_CUDA(ctx, d_pi.upload(pi.data(), npair, s)); CPN_CUDA(ctx, d_pj.upload(pj.data(), npair, s));
        if (matched) CPN_CUDA(ctx, d_js.upload((const int64_t*)labelings, P, s));
        else CPN_CUDA(ctx, d_lab.upload((const int32_t*)labelings, P * n1, s));
        std::vector<int64_t> cp, so;
        for (int64_t r0 = 0; r0 < R; r0 += chunk) {
            const int64_t Rc = std::min(chunk, R - r0), Tm = Rc * S;
            cp.resize(Rc + 1); so.resize(Rc + 1);
            for (int64_t i = 0; i <= Rc; ++i) { cp[i] = cpg_off[r0 + i] - cpg_off[r0]; so[i] = sub_off[r0 + i] - sub_off[r0]; }
            const int64_t n0 = cpg_off[r0], k0 = sub_off[r0], sumN = cp[Rc], sumK = so[Rc];
            CPN_CUDA(ctx, ctx->stage_arena.reset());
            CpnArenaScope scope(&ctx->stage_arena);
            DevIn<double> i_phi, i_rho, i_dn;
            DevIn<int64_t> i_sp, i_sq;
            DevOut<double> o_T, o_p, o_tc;
            DevOut<int64_t> o_cnt;
            DevBuf<double> logZ, lg, e1, e2, mml, nme, cmd;
            DevBuf<int64_t> d_cpg, d_so, reg_of, ex_off, k_off, pa, pb, cmd_off;
            CPN_CUDA(ctx, i_phi.bind(phi + 3 * r0 * S, 3 * Tm, on_device, s));
            CPN_CUDA(ctx, d_cpg.upload(cp.data(), Rc + 1, s)); CPN_CUDA(ctx, d_so.upload(so.data(), Rc + 1, s));
            CPN_CUDA(ctx, i_rho.bind(rho_n + n0, sumN, on_device, s)); CPN_CUDA(ctx, i_dn.bind(d_n + (n0 - r0), sumN - Rc, on_device, s));
            CPN_CUDA(ctx, i_sp.bind(sub_p + k0, sumK, on_device, s)); CPN_CUDA(ctx, i_sq.bind(sub_q + k0, sumK, on_device, s));
            CPN_CUDA(ctx, o_T.bind(T + nstat * k0, nstat * sumK, on_device, s)); CPN_CUDA(ctx, o_cnt.bind(cnt + nstat * k0, nstat * sumK, on_device, s));
            CPN_CUDA(ctx, o_p.bind(pval + nstat * k0, nstat * sumK, on_device, s));
            if (matched) CPN_CUDA(ctx, o_tc.bind(tcmd + k0, sumK, on_device, s));
            CPN_CUDA(ctx, reg_of.alloc(Tm, s)); CPN_CUDA(ctx, ex_off.alloc(Tm + 1, s)); CPN_CUDA(ctx, k_off.alloc(Tm + 1, s));
            CPN_CUDA(ctx, logZ.alloc(Tm, s)); CPN_CUDA(ctx, lg.alloc(4 * S * sumK, s)); CPN_CUDA(ctx, e1.alloc(S * sumK, s));
            CPN_CUDA(ctx, e2.alloc(S * sumK, s)); CPN_CUDA(ctx, mml.alloc(S * sumK, s)); CPN_CUDA(ctx, nme.alloc(S * sumK, s));
            const int64_t ncmd = matched ? (int64_t)n * sumK : (int64_t)S * S * sumK;
            CPN_CUDA(ctx, cmd.alloc(ncmd, s));
            CPN_CUDA(ctx, pa.alloc(Rc * npair, s)); CPN_CUDA(ctx, pb.alloc(Rc * npair, s)); CPN_CUDA(ctx, cmd_off.alloc(Rc * npair, s));
            CPN_CUDA(ctx, cudaMemsetAsync(cmd.p, 0xFF, ncmd * sizeof(double), s));        // NaN outside the pairs that exist
            { KTimer kt(ctx, CPN_K_PREP);
              k_grp_models<<<blocks_for(Tm + 1, 256), 256, 0, s>>>(Tm, S, d_cpg.p, d_so.p, reg_of.p, ex_off.p, k_off.p); }
            { KTimer kt(ctx, CPN_K_PREP);
              k_grp_pairs<<<blocks_for(Rc * npair, 256), 256, 0, s>>>(Rc, S, npair, d_pi.p, d_pj.p, matched, d_so.p, pa.p, pb.p, cmd_off.p); }
            // get_stat_sums! of every sample's model, comp_cmd of every pair, then the labelings
            bool ordered = false;
            if (lane_ok) {
                if (on_device) { int rc0 = cpn_subregions_ordered_device(ctx, Rc, d_so.p, i_sp.p, i_sq.p, d_cpg.p, &ordered); if (rc0) return rc0; }
                else ordered = cpn_subregions_ordered(Rc, so.data(), sub_p + k0, sub_q + k0, cp.data());
            }
            int rc = cpn_summaries_and_cmd(ctx, ordered, Tm, i_phi.p, reg_of.p, ex_off.p, k_off.p, S * sumN, S * sumK, Rc, sumN, d_cpg.p, i_rho.p, i_dn.p,
                                           d_so.p, i_sp.p, i_sq.p, Nmax, Rc * npair, pa.p, pb.p, cmd_off.p, logZ.p, lg.p, e1.p, e2.p, mml.p, nme.p, cmd.p);
            if (rc) return rc;
            if (matched) {
                { KTimer kt(ctx, CPN_K_SWEEP); k_mat_tcmd<<<(unsigned)Rc, 32, 0, s>>>(Rc, n, d_so.p, cmd.p, o_tc.p); }
                { KTimer kt(ctx, CPN_K_SWEEP);
                  k_mat_sweep<<<(unsigned)Rc, SW_THREADS, smem, s>>>(n, Kmax, d_so.p, mml.p, mml.p, nme.p, nme.p, d_js.p, P, exact, o_T.p, o_cnt.p, o_p.p,
                                                                     S, n); }
                CPN_CUDA(ctx, o_tc.finish(s));
            } else {
                KTimer kt(ctx, CPN_K_SWEEP);
                kern<<<(unsigned)Rc, SW_THREADS, smem, s>>>(n1, n2, Kmax, d_so.p, mml.p, nme.p, cmd.p, d_lab.p, P, exact, o_T.p, o_cnt.p, o_p.p);
            }
            CPN_CUDA(ctx, cudaGetLastError());
            CPN_CUDA(ctx, o_T.finish(s)); CPN_CUDA(ctx, o_cnt.finish(s)); CPN_CUDA(ctx, o_p.finish(s));
            CPN_CUDA(ctx, cpn_wait_stream(ctx, s));                       // the chunk's workspace is reused by the next one
            launches += ctx->last_launches;
            ctx->last_launches = 0;
        }
        CPN_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
        CPN_CUDA(ctx, cpn_wait_stream(ctx, s));
    }
    ctx->last_launches = launches;
    float ms = 0;
    CPN_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    if (ctx->profile) cpn_profile_collect(ctx);
    return CPN_OK;
}

int cpn_grp_test_batch(cpn_ctx* ctx, int64_t R, int32_t n1, int32_t n2, int32_t matched, const double* phi, const int64_t* cpg_off,
                       const double* rho_n, const double* d_n, const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q,
                       const void* labelings, int64_t P, int32_t exact, double* T, int64_t* cnt, double* pval, double* tcmd) {
    if (!ctx) return CPN_ERR_ARG;
    if (ctx->devs.size() <= 1 || R < 2)
        return grp_test_impl(ctx, false, R, n1, n2, matched, phi, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, labelings, P, exact, T, cnt, pval, tcmd);
    // cpn_create_multi: equal contiguous blocks of regions (a region's cost is its CpG count, nearly uniform), one host thread per device
    CPN_ARG(ctx, cpg_off && sub_off, "null pointer");
    const int nd = (int)ctx->devs.size(), S = n1 + n2, nstat = matched ? 2 : 3;
    std::vector<int> rc(nd, CPN_OK);
    std::vector<std::thread> th;
    for (int d = 0; d < nd; ++d) {
        th.emplace_back([&, d]() {
            cpn_ctx* c = ctx->devs[d];
            const int64_t r0 = R * d / nd, Rd = R * (d + 1) / nd - r0;
            c->last_ms = 0; c->last_launches = 0;
            if (Rd <= 0) return;
            if (d > 0) cpn_set_profile(c, ctx->profile ? 1 : 0);
            std::vector<int64_t> cp(Rd + 1), so(Rd + 1);
            for (int64_t i = 0; i <= Rd; ++i) { cp[i] = cpg_off[r0 + i] - cpg_off[r0]; so[i] = sub_off[r0 + i] - sub_off[r0]; }
            const int64_t n0 = cpg_off[r0], k0 = sub_off[r0];
            rc[d] = grp_test_impl(c, false, Rd, n1, n2, matched, phi + 3 * S * r0, cp.data(), rho_n + n0, d_n + (n0 - r0), so.data(), sub_p + k0, sub_q + k0,
                                  labelings, P, exact, T + nstat * k0, cnt + nstat * k0, pval + nstat * k0, tcmd ? tcmd + k0 : nullptr);
        });
    }
    for (auto& t : th) t.join();
    cudaSetDevice(ctx->device);
    int ret = CPN_OK;
    double ms = 0;
    int64_t launches = 0;
    for (int d = 0; d < nd; ++d) {
        cpn_ctx* c = ctx->devs[d];
        if (rc[d] != CPN_OK && ret == CPN_OK) { ret = rc[d]; if (d > 0) ctx->err = c->err; }
        ms = std::max(ms, c->last_ms); launches += c->last_launches;
        if (d > 0)
            for (int k = 0; k < CPN_K_NCLASS; ++k) {
                ctx->prof_ms[k] += c->prof_ms[k]; ctx->prof_launches[k] += c->prof_launches[k];
                c->prof_ms[k] = 0; c->prof_launches[k] = 0;
            }
    }
    ctx->last_ms = ms; ctx->last_launches = launches;
    return ret;
}
int cpn_grp_test_batch_device(cpn_ctx* ctx, int64_t R, int32_t n1, int32_t n2, int32_t matched, const double* phi, const int64_t* cpg_off,
                              const double* rho_n, const double* d_n, const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q,
                              const void* labelings, int64_t P, int32_t exact, double* T, int64_t* cnt, double* pval, double* tcmd) {
    if (!ctx) return CPN_ERR_ARG;
    return grp_test_impl(ctx, true, R, n1, n2, matched, phi, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, labelings, P, exact, T, cnt, pval, tcmd);
}

}  // extern "C"
