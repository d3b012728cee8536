// cpn_common.cuh — context, error handling and small device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/cpelnano_b200.h"

// kernel classes timed by the optional per-launch profiler (cpn_set_profile)
enum { CPN_K_PREP = 0, CPN_K_ESTEP = 1, CPN_K_MSTEP = 2, CPN_K_SCORE = 3, CPN_K_PICK = 4, CPN_K_STATSUM = 5, CPN_K_CMD = 6, CPN_K_SWEEP = 7,
       CPN_K_NCLASS = 8 };

// Grow-only device workspace with stack discipline.  The fit path allocates the same sequence of (large) buffers on every
// call; taking them from the stream-ordered pool made the driver defragment the pool under multi-stream use (stalls of
// hundreds of ms), so they come from one persistent block instead.  A request that does not fit falls back to the pool and
// raises the size the block is re-created with at the next reset().
struct CpnArena {
    char* base = nullptr;
    size_t cap = 0, vtop = 0, high = 0;
    struct Spill { void* p; size_t voff, bytes; cudaStream_t s; };
    std::vector<Spill> spill;
    static size_t round_up(size_t b) { return (b + 511) & ~(size_t)511; }
    cudaError_t alloc(void** out, size_t bytes, cudaStream_t s) {
        size_t b = round_up(bytes ? bytes : 1), voff = vtop;
        vtop += b;
        if (vtop > high) high = vtop;
        if (voff + b <= cap) { *out = base + voff; return cudaSuccess; }
        cudaError_t e = cudaMallocAsync(out, b, s);
        if (e == cudaSuccess) spill.push_back({*out, voff, b, s});
        return e;
    }
    void release(void* p, size_t bytes) {
        size_t b = round_up(bytes ? bytes : 1);
        if (base && (char*)p >= base && (char*)p < base + cap) {
            size_t voff = (size_t)((char*)p - base);
            if (voff + b == vtop) vtop = voff;          // top of the stack: reusable right away (same stream)
            return;
        }
        for (size_t i = 0; i < spill.size(); ++i)
            if (spill[i].p == p) {
                if (spill[i].voff + spill[i].bytes == vtop) vtop = spill[i].voff;
                cudaFreeAsync(p, spill[i].s);
                spill.erase(spill.begin() + i);
                return;
            }
    }
    // Call only when no stream still uses memory handed out earlier.
    cudaError_t reset() {
        for (auto& sp : spill) cudaFreeAsync(sp.p, sp.s);
        spill.clear();
        vtop = 0;
        if (high > cap) {
            if (base) cudaFree(base);
            base = nullptr; cap = 0;
            size_t want = (high + high / 8 + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
            cudaError_t e = cudaMalloc((void**)&base, want);
            if (e != cudaSuccess) { cudaGetLastError(); return cudaSuccess; }   // keep spilling to the pool
            cap = want;
        }
        return cudaSuccess;
    }
    void destroy() {
        for (auto& sp : spill) cudaFreeAsync(sp.p, sp.s);
        spill.clear();
        if (base) cudaFree(base);
        base = nullptr; cap = vtop = high = 0;
    }
};
// The arena DevBuf allocations of the calling thread come from (null: the stream-ordered pool).
inline CpnArena*& cpn_tl_arena() { static thread_local CpnArena* a = nullptr; return a; }
struct CpnArenaScope {
    CpnArena* prev;
    explicit CpnArenaScope(CpnArena* a) : prev(cpn_tl_arena()) { cpn_tl_arena() = a; }
    ~CpnArenaScope() { cpn_tl_arena() = prev; }
};

struct cpn_ctx {
    int device = 0;
    int sm_count = 148;
    bool ep_attr_set = false;                // k_estep_pair's dynamic shared-memory limit raised on this device
    int epw_blocks = 1;                      // resident k_estep_pair_ws blocks per SM (occupancy query, ep_prepare)
    int estep_blocks = 0;                    // resident k_estep blocks per SM (occupancy query, cached by the first E-step launch)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_wait = nullptr;           // blocking-sync event behind cpn_wait_stream
    std::string err;
    double last_ms = 0.0;
    int64_t last_launches = 0;
    // algorithmic work of the last fit call, for roofline accounting: Σ E-step passes·M·L, Σ derivative sweeps (or log Z
    // evaluations)·L, Σ score passes·M·L, EM instances
    double last_work[4] = {0, 0, 0, 0};
    uint64_t seed = 0;                       // cpn_set_seed: key of the counter-based start / permutation stream (cpn_philox.h)
    // profiler: when enabled every launch is bracketed by events on `stream`
    bool profile = false;
    std::vector<cudaEvent_t> pev;            // event pool
    std::vector<int> pev_class;              // class of the launch between pev[2i] and pev[2i+1]
    size_t pev_used = 0;
    double prof_ms[CPN_K_NCLASS] = {0};
    int64_t prof_launches[CPN_K_NCLASS] = {0};
    // worker sub-contexts (own stream/events) used by the host-pointer entry points to run chunks of a batch concurrently
    std::vector<cpn_ctx*> children;
    // cpn_create_multi: one context per device, devs[0] == this.  Host-pointer entry points shard the regions of a call over
    // them (one host thread per device); region_base = position of a shard's first region in the caller's batch (seeded starts
    // and permutations are keyed by the global position, so sharding does not change a result)
    std::vector<cpn_ctx*> devs;
    int64_t region_base = 0;
    // EM loop: pinned mailbox for the per-iteration active counts and the events that say when each has arrived
    int* h_counts = nullptr;
    int h_counts_cap = 0;
    std::vector<cudaEvent_t> it_ev;
    // persistent workspaces of the fit path: EM state / outputs, staged inputs, and the staging slots of a chunked host call
    CpnArena arena, stage_arena;
    std::vector<CpnArena*> slot_arenas;
};

// Host waits sleep instead of spinning (cudaEventBlockingSync): a host-pointer call keeps a staging thread and two workers
// waiting on the GPU most of the time, and with one process per GPU on a shared host (8 ranks × 3 threads on 16 cores) spinning
// waiters take the cores the launching threads need.  The wake-up latency (tens of µs) is off the critical path: the EM loop
// runs two iterations ahead of the device.  CPN_SPIN_WAIT=1 restores cudaStreamSynchronize.
inline bool cpn_spin_wait() {
    static const bool spin = getenv("CPN_SPIN_WAIT") != nullptr && atoi(getenv("CPN_SPIN_WAIT")) != 0;
    return spin;
}
inline cudaError_t cpn_wait_stream(cpn_ctx* ctx, cudaStream_t s) {
    if (cpn_spin_wait()) return cudaStreamSynchronize(s);
    cudaError_t e = cudaSuccess;
    if (!ctx->ev_wait) e = cudaEventCreateWithFlags(&ctx->ev_wait, cudaEventBlockingSync | cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_wait, s);
    if (e == cudaSuccess) e = cudaEventSynchronize(ctx->ev_wait);
    return e;
}

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
    CpnArena* ar = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) {
            if (ar) ar->release(p, (n ? n : 1) * sizeof(T));
            else cudaFreeAsync(p, s);
        }
        p = nullptr; n = 0; ar = nullptr;
    }
    cudaError_t alloc(size_t count, cudaStream_t stream) {
        release();
        s = stream; n = count;
        if (count == 0) count = 1;
        ar = cpn_tl_arena();
        if (ar) return ar->alloc((void**)&p, count * sizeof(T), stream);
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
                         const int64_t* sub_q, int64_t Nmax, double* logZ, double* ex /*null: not written*/, double* exx, double* log_g,
                         double* e_log_g1, double* e_log_g2, double* mml, double* nme, double* msum /*[k_off[T]][5] or null*/);
// Host check: sub-regions of every region ascending, disjoint and at most 32 (what get_nls_reg_info! produces) — the condition
// for the lane-per-pair CMD kernel.
bool cpn_subregions_ordered(int64_t R, const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q, const int64_t* cpg_off);
int cpn_subregions_ordered_device(cpn_ctx* ctx, int64_t R, const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q, const int64_t* cpg_off,
                                  bool* ordered);
// get_stat_sums! of T models + comp_cmd of P pairs of them (device pointers; scratch from the calling thread's arena / pool).
int cpn_summaries_and_cmd(cpn_ctx* ctx, bool ordered, int64_t T, const double* phi, const int64_t* reg_of, const int64_t* ex_off,
                          const int64_t* k_off, int64_t sN, int64_t sK, int64_t R, int64_t sumN, const int64_t* cpg_off, const double* rho_n,
                          const double* d_n, const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q, int64_t Nmax, int64_t P,
                          const int64_t* pair_a, const int64_t* pair_b, const int64_t* cmd_off, double* logZ, double* log_g, double* e_log_g1,
                          double* e_log_g2, double* mml, double* nme, double* cmd);

// Launch-only entry of the CMD kernel (all pointers on the device); defined in cpn_summaries.cu.
int cpn_cmd_launch(cpn_ctx* ctx, int64_t P, const int64_t* pair_a, const int64_t* pair_b, const double* phi_tbl, const int64_t* reg_of,
                   const int64_t* ex_off, const double* ex_tbl, const double* exx_tbl, const int64_t* nme_off, const double* nme_tbl,
                   const int64_t* cpg_off, const double* rho_n, const double* d_n, const int64_t* sub_off, const int64_t* sub_p,
                   const int64_t* sub_q, int64_t Nmax, const int64_t* cmd_off, double* cmd);

// Launch-only entry of the two-sample p-value counting kernel (all pointers on the device); defined in cpn_sweeps.cu.
int cpn_two_samp_pvals_launch(cpn_ctx* ctx, int64_t R, int Kmax, const int64_t* tbl_off, const int64_t* perm_off, const int64_t* nul_off,
                              const int32_t* exact_r, const double* T_obs, const double* T_null, int64_t* cnt, int64_t* n_valid, double* pval);

struct LaunchCounter {
    cpn_ctx* ctx;
    explicit LaunchCounter(cpn_ctx* c) : ctx(c) { ctx->last_launches = 0; for (double& w : ctx->last_work) w = 0; }
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
// exp(x) for x <= 0 (the only exponentials on the hot path are e^{-2|α|}, e^{-2|β|}).  64-entry table of 2^(j/64),
// degree-5 polynomial on |r| <= ln2/128, exponent patched in integer arithmetic: 10 FP64 instructions instead of the
// ~21 of the general-purpose exp(); ≤ 1.5 ulp; results below 2^-1021 flush to 0 (they multiply messages of size O(1)).
__device__ const double cpn_exp_tbl[64] = {1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284, 1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199, 1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418, 1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812, 1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687, 1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783, 1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303, 1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112, 1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647, 1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384, 1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267, 1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364, 1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062, 1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989, 1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656, 1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951};
__device__ __forceinline__ double cpn_exp_neg(double x) {
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52: adds round-to-nearest-integer in the low word
    double kd = fma(x, 92.332482616893656, MAGIC);            // 64 / ln 2
    int ki = __double2loint(kd);
    kd -= MAGIC;
    double r = fma(kd, -0.01083042469326756, x);              // ln2/64, high part (21 trailing zero bits: kd*hi is exact)
    r = fma(kd, -2.9815858269852933e-12, r);                  // ln2/64, low part
    double q = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
    q = fma(q, r, 1.6666666666666666e-01);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = q * r;                                                // e^r - 1
    double T = __ldg(&cpn_exp_tbl[ki & 63]);
    double y = fma(T, q, T);
    int e = ki >> 6;
    int hi = __double2hiint(y) + (e << 20);
    y = __hiloint2double(hi, __double2loint(y));
    // x < -746, -Inf → 0; NaN (arguments are built as -2·|·|, so a NaN carries the sign bit) → NaN.  Integer tests only.
    unsigned uhi = (unsigned)__double2hiint(x);
    if (uhi > 0xC0875000u) {
        bool isnan_ = ((uhi & 0x7ff00000u) == 0x7ff00000u) && (((uhi & 0x000fffffu) | (unsigned)__double2loint(x)) != 0u);
        return isnan_ ? x : 0.0;
    }
    return (e < -1021) ? 0.0 : y;
}
// cpn_exp_neg for arguments that are known not to be NaN-carrying results the caller still needs (the M-step sweep: a non-finite
// trial point is caught by the solver's own checks).  The argument's high word is clamped at −708 with one integer minimum
// (for negative doubles a larger magnitude is a larger unsigned high word; −Inf and sign-carrying NaNs clamp too), so the exponent
// patch never leaves the normal range and the range tests of cpn_exp_neg (7 instructions, a branch) disappear.  Same bits as
// cpn_exp_neg for −708 ≤ x ≤ 0; below that e^{−708} ≈ 3·10⁻³⁰⁸ instead of 0.
__device__ __forceinline__ double cpn_exp_neg_clamped(double x) {
    const double MAGIC = 6755399441055744.0;
    x = __hiloint2double((int)min((unsigned)__double2hiint(x), 0xC0862000u), __double2loint(x));
    double kd = fma(x, 92.332482616893656, MAGIC);
    int ki = __double2loint(kd);
    kd -= MAGIC;
    double r = fma(kd, -0.01083042469326756, x);
    r = fma(kd, -2.9815858269852933e-12, r);
    double q = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
    q = fma(q, r, 1.6666666666666666e-01);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = q * r;
    double T = __ldg(&cpn_exp_tbl[ki & 63]);
    double y = fma(T, q, T);
    return __hiloint2double(__double2hiint(y) + ((ki >> 6) << 20), __double2loint(y));
}
__device__ __forceinline__ double cpn_nan() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ double cpn_inf() { return __longlong_as_double(0x7ff0000000000000LL); }
#endif
