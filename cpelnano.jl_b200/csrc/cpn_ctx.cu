// cpn_ctx.cu — context lifetime, error text, timing accessors, FP64 peak probe, and the host-only
// integer routine get_nls_reg_info! (src/genome_analysis.jl:472-538).
#include <cmath>
#include <vector>

#include "cpn_common.cuh"

namespace {

// 8 independent DFMA chains per thread: enough ILP to saturate the FP64 pipe at full occupancy.
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace

extern "C" {

const char* cpn_version(void) { return "cpelnano_b200 0.1 (sm_100a)"; }

int cpn_create(int device_id, cpn_ctx** out) {
    if (!out) return CPN_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device_id < 0 || device_id >= n) return CPN_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) return CPN_ERR_CUDA;
    if (prop.major != 10) return CPN_ERR_CUDA;      // kernels are built for sm_100a only; no fallback
    if (cudaSetDevice(device_id) != cudaSuccess) return CPN_ERR_CUDA;
    cpn_ctx* ctx = new cpn_ctx();
    ctx->device = device_id;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&ctx->ev0) != cudaSuccess ||
        cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        delete ctx;
        return CPN_ERR_CUDA;
    }
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = ctx;
    return CPN_OK;
}

void cpn_destroy(cpn_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_wait) cudaEventDestroy(ctx->ev_wait);
    for (cudaEvent_t e : ctx->pev) cudaEventDestroy(e);
    for (cpn_ctx* c : ctx->children) cpn_destroy(c);
    for (cudaEvent_t e : ctx->it_ev) cudaEventDestroy(e);
    ctx->arena.destroy();
    ctx->stage_arena.destroy();
    for (CpnArena* a : ctx->slot_arenas) { a->destroy(); delete a; }
    if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
    delete ctx;
}

const char* cpn_last_error(const cpn_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
double cpn_last_device_ms(const cpn_ctx* ctx) { return ctx ? ctx->last_ms : 0.0; }
int64_t cpn_last_launches(const cpn_ctx* ctx) { return ctx ? ctx->last_launches : 0; }

int cpn_set_profile(cpn_ctx* ctx, int enable) {
    if (!ctx) return CPN_ERR_ARG;
    ctx->profile = enable != 0;
    for (int i = 0; i < CPN_K_NCLASS; ++i) { ctx->prof_ms[i] = 0.0; ctx->prof_launches[i] = 0; }
    ctx->pev_used = 0;
    return CPN_OK;
}
int cpn_get_profile(const cpn_ctx* ctx, double* ms, int64_t* launches) {
    if (!ctx || !ms || !launches) return CPN_ERR_ARG;
    for (int i = 0; i < CPN_K_NCLASS; ++i) { ms[i] = ctx->prof_ms[i]; launches[i] = ctx->prof_launches[i]; }
    return CPN_OK;
}

int cpn_measure_fp64_peak(cpn_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return CPN_ERR_ARG;
    CPN_CUDA(ctx, cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 20000;
    DevBuf<double> out;
    CPN_CUDA(ctx, out.alloc((size_t)blocks * threads, ctx->stream));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CPN_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        k_dfma_peak<<<blocks, threads, 0, ctx->stream>>>(out.p, iters, 1.0 + rep);
        CPN_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CPN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        CPN_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    *tflops = best;
    return CPN_OK;
}

// get_nls_reg_info!  src/genome_analysis.jl:472-538 — K = ceil(size/max_size_subreg) sub-regions, the
// remainder spread one base at a time from the left, closed intervals sharing end points, CpG n in
// sub-region i iff bnds[i] <= pos < bnds[i+1]; the last sub-region always ends at N; empty = 0.
int64_t cpn_nls_reg_info(int64_t chrst, int64_t chrend, int64_t max_size_subreg, const int64_t* cpg_pos, int64_t N, int64_t* sub_st,
                         int64_t* sub_end, int64_t* sub_p, int64_t* sub_q) {
    if (max_size_subreg <= 0 || chrend < chrst || N < 0 || (N > 0 && !cpg_pos)) return -1;
    int64_t size = chrend - chrst + 1;
    int64_t k = (int64_t)std::ceil((double)size / (double)max_size_subreg);
    if (k < 1) return -1;
    if (!sub_st) return k;
    if (!sub_end || !sub_p || !sub_q) return -1;
    int64_t base = (int64_t)std::floor((double)size / (double)k);
    int64_t rem = size % k;
    int64_t lo = chrst;                      // left boundary of sub-region j
    int64_t acc = 0;
    int64_t n0 = 0;                          // first CpG index (0-based) not yet assigned
    for (int64_t j = 0; j < k; ++j) {
        acc += base + (j < rem ? 1 : 0);
        int64_t hi = chrst - 1 + acc;        // right boundary (shared with the next sub-region)
        sub_st[j] = lo; sub_end[j] = hi;
        // positions are sorted: skip those left of lo, then take those < hi
        while (n0 < N && cpg_pos[n0] < lo) ++n0;
        int64_t n1 = n0;
        while (n1 < N && cpg_pos[n1] < hi) ++n1;
        int64_t fst = (n1 > n0) ? n0 + 1 : 0, lst = (n1 > n0) ? n1 : 0;
        if (j < k - 1) { sub_p[j] = fst; sub_q[j] = lst; }
        else { sub_p[j] = fst; sub_q[j] = N; }
        n0 = n1;
        lo = hi;
    }
    return k;
}

}  // extern "C"
