// cpn_ctx.cu — context lifetime, error text, timing accessors, FP64 peak probe, and the host-only
// integer routine get_nls_reg_info! (src/genome_analysis.jl:472-538).
#include <cmath>
#include <vector>

#include "cpn_common.cuh"
#include "cpn_philox.h"

namespace {

// 8 independent DFMA chains per thread, 128 threads per block, 8 blocks per SM: the configuration of tools/micro/fp64_mix.cu,
// which reads 36.7 TFLOP/s on a B200 (nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2).  The loop body is fully unrolled
// DFMAs; the launch is long (>= 10 ms) so that clock ramp-up and the launch's head and tail do not weigh on the figure
// (the round-1 probe ran 2.6 ms launches and under-read by about 6 %).
__global__ void __launch_bounds__(128, 8) k_dfma_peak(double* out, int iters, double seed) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + k + tid * 1e-9;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = fma(a[k], m, c);
    }
    double z = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) z += a[k];
    if (z == 123.456) out[tid] = z;          // never true: keeps the chains alive without a store on the timed path
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

int cpn_create_multi(const int* device_ids, int n_dev, cpn_ctx** out) {
    if (!out || !device_ids || n_dev < 1) return CPN_ERR_ARG;
    *out = nullptr;
    for (int i = 0; i < n_dev; ++i)
        for (int j = 0; j < i; ++j)
            if (device_ids[i] == device_ids[j]) return CPN_ERR_ARG;
    cpn_ctx* head = nullptr;
    int rc = cpn_create(device_ids[0], &head);
    if (rc != CPN_OK) return rc;
    head->devs.push_back(head);
    for (int i = 1; i < n_dev; ++i) {
        cpn_ctx* c = nullptr;
        rc = cpn_create(device_ids[i], &c);
        if (rc != CPN_OK) { cpn_destroy(head); return rc; }
        head->devs.push_back(c);
    }
    cudaSetDevice(device_ids[0]);
    *out = head;
    return CPN_OK;
}
int cpn_num_devices(const cpn_ctx* ctx) { return ctx ? (ctx->devs.empty() ? 1 : (int)ctx->devs.size()) : 0; }

void cpn_destroy(cpn_ctx* ctx) {
    if (!ctx) return;
    for (size_t i = 1; i < ctx->devs.size(); ++i) cpn_destroy(ctx->devs[i]);
    ctx->devs.clear();
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
int cpn_last_work(const cpn_ctx* ctx, double* work) {
    if (!ctx || !work) return CPN_ERR_ARG;
    for (int k = 0; k < 4; ++k) work[k] = ctx->last_work[k];
    return CPN_OK;
}

int cpn_set_seed(cpn_ctx* ctx, uint64_t seed) {
    if (!ctx) return CPN_ERR_ARG;
    ctx->seed = seed;
    return CPN_OK;
}

// Host-side evaluation of the same stream (no GPU needed): what the device generates for phi0 == NULL / perm_ids == NULL.
int cpn_rng_phi0(uint64_t seed, int64_t region, int64_t perm, int32_t half, int32_t n_init, double* phi0) {
    if (!phi0 || n_init < 0 || region < 0 || half < 0 || half > 1) return CPN_ERR_ARG;
    const uint32_t p = perm < 0 ? CPN_RNG_NO_PERM : (uint32_t)perm;
    for (int32_t s = 0; s < n_init; ++s) cpn_gen_phi0(seed, (uint32_t)region, p, (uint32_t)half, (uint32_t)s, phi0 + 3 * s);
    return CPN_OK;
}
int cpn_rng_perm(uint64_t seed, int64_t region, int64_t perm, int32_t n, int32_t n1, int32_t* ids) {
    if (!ids || n < 1 || n1 < 0 || n1 > n || region < 0 || perm < 0) return CPN_ERR_ARG;
    int need = n1, k = 0;
    CpnPhilox4 blk{};
    for (int q = 0; q < n; ++q) {
        if ((q & 3) == 0) blk = cpn_philox4x32_10(seed, (uint32_t)region, (uint32_t)perm, (uint32_t)(q >> 2), 1u);
        if ((int)cpn_mulhi32(blk.x[q & 3], (uint32_t)(n - q)) < need) { ids[k++] = q + 1; --need; }
    }
    return CPN_OK;
}

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
    const int blocks = ctx->sm_count * 8, threads = 128, iters = 200000;
    DevBuf<double> out;
    CPN_CUDA(ctx, out.alloc((size_t)blocks * threads, ctx->stream));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {                      // rep 0 warms the clocks up; best of the next four
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
