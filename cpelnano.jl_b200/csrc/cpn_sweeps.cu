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
#include <cmath>

#include "cpn_common.cuh"

namespace {

constexpr int SW_THREADS = 256;

// sum(vector of K-vectors)/n with Julia's left-to-right `+` (grp_perm_tests.jl:121-125)
__device__ __forceinline__ double mean_ids(const double* tab, int K, int k, const int* ids, int n) {
    double s = tab[(ids[0] - 1) * K + k];
    for (int i = 1; i < n; ++i) s = __dadd_rn(s, tab[(ids[i] - 1) * K + k]);
    return __ddiv_rn(s, (double)n);
}

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
    int* s_cnt = (int*)(s_T + 3 * Kmax);         // [3][K]
    for (int i = threadIdx.x; i < S * K; i += blockDim.x) { s_mml[i] = mml[k0 * S + i]; s_nme[i] = nme[k0 * S + i]; }
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
    int g1[32], g2[32];
    for (int64_t pi = threadIdx.x; pi < P; pi += blockDim.x) {
        const int32_t* lab = labelings + pi * n1;
        for (int i = 0; i < n1; ++i) g1[i] = lab[i];
        int a = 0, c = 0;
        for (int s = 1; s <= S; ++s) { if (a < n1 && g1[a] == s) ++a; else g2[c++] = s; }
        for (int k = 0; k < K; ++k) {
            double tobs = s_T[k];
            if (tobs != tobs) continue;                                          // isnan(tmml_obs[k]) && continue (:260)
            double tm = __dsub_rn(mean_ids(s_mml, K, k, g1, n1), mean_ids(s_mml, K, k, g2, n2));
            double tn = __dsub_rn(mean_ids(s_nme, K, k, g1, n1), mean_ids(s_nme, K, k, g2, n2));
            double tc = 0.0;                                                     // comp_unmat_perm_stat_cmd :86-107
            for (int i = 0; i < n1; ++i)
                for (int j = 0; j < n2; ++j) {
                    int ii = g1[i] - 1, jj = g2[j] - 1;
                    int lo = min(ii, jj), hi = max(ii, jj);
                    tc = __dadd_rn(tc, s_cmd[(lo * S + hi) * K + k]);
                }
            tc = __ddiv_rn(tc, (double)(n1 * n2));
            if (fabs(tm) >= fabs(tobs)) atomicAdd(&s_cnt[k], 1);                 // :268-270
            if (fabs(tn) >= fabs(s_T[K + k])) atomicAdd(&s_cnt[K + k], 1);
            if (tc >= s_T[2 * K + k]) atomicAdd(&s_cnt[2 * K + k], 1);
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

__global__ void __launch_bounds__(SW_THREADS) k_mat_sweep(int n, int Kmax, const int64_t* __restrict__ tbl_off, const double* __restrict__ mml1,
                                                           const double* __restrict__ mml2, const double* __restrict__ nme1,
                                                           const double* __restrict__ nme2, const int64_t* __restrict__ js, int64_t P,
                                                           int exact, double* __restrict__ T, int64_t* __restrict__ cnt,
                                                           double* __restrict__ pval) {
    extern __shared__ double sm[];
    int64_t r = blockIdx.x;
    int64_t k0 = tbl_off[r];
    int K = (int)(tbl_off[r + 1] - k0);
    double* s_dm = sm;                          // [n][K] pair differences (comp_mat_diff_mml :343)
    double* s_dn = s_dm + n * Kmax;
    double* s_T = s_dn + n * Kmax;              // [2][K]
    int* s_cnt = (int*)(s_T + 2 * Kmax);
    for (int i = threadIdx.x; i < n * K; i += blockDim.x) {
        s_dm[i] = __dsub_rn(mml1[k0 * n + i], mml2[k0 * n + i]);
        s_dn[i] = __dsub_rn(nme1[k0 * n + i], nme2[k0 * n + i]);
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
            a = __ddiv_rn(a, (double)n); b = __ddiv_rn(b, (double)n);
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

__global__ void __launch_bounds__(SW_THREADS) k_two_samp_pvals(int Kmax, const int64_t* __restrict__ tbl_off, int64_t P,
                                                                const double* __restrict__ T_obs, const double* __restrict__ T_null, int exact,
                                                                int64_t* __restrict__ cnt, int64_t* __restrict__ n_valid, double* __restrict__ pval) {
    extern __shared__ double sm[];
    int64_t r = blockIdx.x;
    int64_t k0 = tbl_off[r];
    int K = (int)(tbl_off[r + 1] - k0);
    double* s_T = sm;                            // [3][K]
    int* s_cnt = (int*)(s_T + 3 * Kmax);         // [3][K]
    int* s_nv = s_cnt + 3 * Kmax;                // [3][K]
    for (int i = threadIdx.x; i < 3 * K; i += blockDim.x) { s_T[i] = T_obs[3 * k0 + i]; s_cnt[i] = 0; s_nv[i] = 0; }
    __syncthreads();
    const double* nul = T_null + 3 * k0 * P;
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

int kmax_of(const int64_t* tbl_off, int64_t R) {
    int64_t k = 1;
    for (int64_t r = 0; r < R; ++r) k = std::max(k, tbl_off[r + 1] - tbl_off[r]);
    return (int)k;
}

}  // namespace

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
    size_t smem = (size_t)(2 * S + S * S + 3) * Kmax * sizeof(double) + (size_t)3 * Kmax * sizeof(int);
    CPN_ARG(ctx, smem <= 200 * 1024, "statistic tables of one region do not fit in shared memory");
    for (int64_t i = 0; i < P * n1; ++i) CPN_ARG(ctx, labelings[i] >= 1 && labelings[i] <= S, "labeling id out of range");
    cudaStream_t s = ctx->stream;
    CPN_CUDA(ctx, cudaFuncSetAttribute(k_unmat_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CPN_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    {
        DevBuf<int64_t> d_off, o_cnt;
        DevBuf<double> d_mml, d_nme, d_cmd, o_T, o_p;
        DevBuf<int32_t> d_lab;
        CPN_CUDA(ctx, d_off.upload(tbl_off, R + 1, s)); CPN_CUDA(ctx, d_mml.upload(mml, sumK * S, s)); CPN_CUDA(ctx, d_nme.upload(nme, sumK * S, s));
        CPN_CUDA(ctx, d_cmd.upload(cmd_tbl, sumK * S * S, s)); CPN_CUDA(ctx, d_lab.upload(labelings, P * n1, s));
        CPN_CUDA(ctx, o_T.alloc(3 * sumK, s)); CPN_CUDA(ctx, o_cnt.alloc(3 * sumK, s)); CPN_CUDA(ctx, o_p.alloc(3 * sumK, s));
        { KTimer kt(ctx, CPN_K_SWEEP); k_unmat_sweep<<<(unsigned)R, SW_THREADS, smem, s>>>(n1, n2, Kmax, d_off.p, d_mml.p, d_nme.p, d_cmd.p, d_lab.p, P, exact, o_T.p, o_cnt.p, o_p.p); }
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
        { KTimer kt(ctx, CPN_K_SWEEP); k_mat_sweep<<<(unsigned)R, SW_THREADS, smem, s>>>(n, Kmax, d_off.p, d_m1.p, d_m2.p, d_n1.p, d_n2.p, d_js.p, P, exact, o_T.p, o_cnt.p, o_p.p); }
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
        { KTimer kt(ctx, CPN_K_SWEEP); k_two_samp_pvals<<<(unsigned)R, SW_THREADS, smem, s>>>(Kmax, d_off.p, P, d_T.p, d_null.p, exact, o_cnt.p, o_nv.p, o_p.p); }
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

}  // extern "C"
