// cpn_common.cuh — context, error handling and small device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/cpelnano_b200.h"

// kernel classes timed by the optional per-launch profiler (cpn_set_profile)
enum { CPN_K_PREP = 0, CPN_K_ESTEP = 1, CPN_K_MSTEP = 2, CPN_K_SCORE = 3, CPN_K_PICK = 4, CPN_K_STATSUM = 5, CPN_K_CMD = 6, CPN_K_SWEEP = 7,
       CPN_K_NCLASS = 8 };

struct cpn_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    double last_ms = 0.0;
    int64_t last_launches = 0;
    // profiler: when enabled every launch is bracketed by events on `stream`
    bool profile = false;
    std::vector<cudaEvent_t> pev;            // event pool
    std::vector<int> pev_class;              // class of the launch between pev[2i] and pev[2i+1]
    size_t pev_used = 0;
    double prof_ms[CPN_K_NCLASS] = {0};
    int64_t prof_launches[CPN_K_NCLASS] = {0};
};

// Brackets one launch with events when profiling is on.  Usage: { KTimer t(ctx, CPN_K_ESTEP); kernel<<<...>>>(...); }
struct KTimer {
    cpn_ctx* ctx;
    size_t slot = 0;
    bool on;
    KTimer(cpn_ctx* c, int cls) : ctx(c), on(c->profile) {
        ctx->last_launches += 1;
        if (!on) return;
        if (ctx->pev_used + 2 > ctx->pev.size()) {
            for (int i = 0; i < 64; ++i) { cudaEvent_t e; cudaEventCreate(&e); ctx->pev.push_back(e); }
        }
        slot = ctx->pev_used;
        ctx->pev_used += 2;
        ctx->pev_class.resize(ctx->pev_used / 2);
        ctx->pev_class[slot / 2] = cls;
        cudaEventRecord(ctx->pev[slot], ctx->stream);
    }
    ~KTimer() { if (on) cudaEventRecord(ctx->pev[slot + 1], ctx->stream); }
};
// Folds the recorded events into prof_ms (call after a stream synchronize).
inline void cpn_profile_collect(cpn_ctx* ctx) {
    for (size_t i = 0; i + 1 < ctx->pev_used; i += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->pev[i], ctx->pev[i + 1]) == cudaSuccess) {
            ctx->prof_ms[ctx->pev_class[i / 2]] += ms;
            ctx->prof_launches[ctx->pev_class[i / 2]] += 1;
        }
    }
    ctx->pev_used = 0;
}

#define CPN_CUDA(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                 \
            return CPN_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define CPN_ARG(ctx, cond, msg)                                                               \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            (ctx)->err = std::string("bad argument: ") + (msg);                               \
            return CPN_ERR_ARG;                                                               \
        }                                                                                     \
    } while (0)

// Stream-ordered device buffer (freed on the same stream at scope exit).
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFreeAsync(p, s);
        p = nullptr; n = 0;
    }
    cudaError_t alloc(size_t count, cudaStream_t stream) {
        release();
        s = stream; n = count;
        if (count == 0) count = 1;
        return cudaMallocAsync((void**)&p, count * sizeof(T), stream);
    }
    cudaError_t upload(const T* host, size_t count, cudaStream_t stream) {
        cudaError_t e = alloc(count, stream);
        if (e != cudaSuccess) return e;
        if (count == 0) return cudaSuccess;
        return cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, stream);
    }
    cudaError_t download(T* host, size_t count, cudaStream_t stream) const {
        if (count == 0 || host == nullptr) return cudaSuccess;
        return cudaMemcpyAsync(host, p, count * sizeof(T), cudaMemcpyDeviceToHost, stream);
    }
};

// A pointer that is either borrowed (already on the device) or owned (uploaded from the host).
template <typename T>
struct DevIn {
    DevBuf<T> own;
    const T* p = nullptr;
    cudaError_t bind(const T* src, size_t count, bool on_device, cudaStream_t s) {
        if (on_device) { p = src; return cudaSuccess; }
        cudaError_t e = own.upload(src, count, s);
        p = own.p;
        return e;
    }
};
template <typename T>
struct DevOut {
    DevBuf<T> own;
    T* p = nullptr;
    T* host = nullptr;
    size_t n = 0;
    cudaError_t bind(T* dst, size_t count, bool on_device, cudaStream_t s) {
        n = count;
        if (dst == nullptr) { host = nullptr; cudaError_t e = own.alloc(count, s); p = own.p; return e; }
        if (on_device) { p = dst; host = nullptr; return cudaSuccess; }
        host = dst;
        cudaError_t e = own.alloc(count, s);
        p = own.p;
        return e;
    }
    cudaError_t finish(cudaStream_t s) const {
        if (host == nullptr) return cudaSuccess;
        return own.download(host, n, s);
    }
};

// Launch-only entry of the summaries kernel (all pointers on the device); defined in cpn_summaries.cu.
int cpn_stat_sums_launch(cpn_ctx* ctx, int64_t T, const double* phi, const int64_t* reg_of, const int64_t* ex_off, const int64_t* k_off,
                         const int64_t* cpg_off, const double* rho_n, const double* d_n, const int64_t* sub_off, const int64_t* sub_p,
                         const int64_t* sub_q, int64_t Nmax, double* logZ, double* ex, double* exx, double* log_g, double* e_log_g1,
                         double* e_log_g2, double* mml, double* nme);

struct LaunchCounter {
    cpn_ctx* ctx;
    explicit LaunchCounter(cpn_ctx* c) : ctx(c) { ctx->last_launches = 0; }
};

// ---- device helpers -------------------------------------------------------------------------------
#ifdef __CUDACC__
// 2^-(unbiased exponent of s) for s > 0 finite; *e receives the unbiased exponent.
__device__ __forceinline__ double cpn_inv_pow2_of(double s, int* e) {
    int be = (__double2hiint(s) >> 20) & 0x7ff;
    *e = be - 1023;
    return __hiloint2double((2046 - be) << 20, 0);
}
__device__ __forceinline__ double cpn_inv_pow2_of(double s) {
    int be = (__double2hiint(s) >> 20) & 0x7ff;
    return __hiloint2double((2046 - be) << 20, 0);
}
__device__ __forceinline__ double cpn_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double cpn_nan() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ double cpn_inf() { return __longlong_as_double(0x7ff0000000000000LL); }
#endif
