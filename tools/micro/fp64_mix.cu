// fp64_mix.cu — what does the FP64 pipe deliver for the E-step's chain step when nothing else is in the way?
// The chain step of k_estep (cpn_em.cu::estep_site: 27 FP64 instructions in three dependent layers, power-of-two rescale every
// second step) runs here on register-resident operands only: no global or shared memory, no tile fill, no reductions.
// Compared with a pure independent-DFMA loop (the peak probe) at the same occupancy.  Build and run:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o fp64_mix tools/micro/fp64_mix.cu && ./fp64_mix
#include <cstdio>
#include <cuda_runtime.h>

struct EState { double fp, fm, gap, gam, gbp, gbm, gcp, gcm; };

__device__ __forceinline__ double inv_pow2_of(double x) {
    int hi = __double2hiint(x);
    int e = ((hi >> 20) & 0x7ff) - 1023;
    return __hiloint2double((1023 - e) << 20, 0);
}

__device__ __forceinline__ void site(EState& s, double wx, double wy, double t, double na, double nb, double sinvd, double r, double sc) {
    double phiP = (wy * r) * sc, phiM = wx * sc;
    double u = s.fp, v = s.fm;
    double Gp = fma(t, v, u), Gm = fma(t, u, v);
    double Hap = fma(t, s.gam, s.gap), Ham = fma(t, s.gap, s.gam);
    double Hbp = fma(t, s.gbm, s.gbp), Hbm = fma(t, s.gbp, s.gbm);
    double Hcp = fma(t, s.gcm, s.gcp), Hcm = fma(t, s.gcp, s.gcm);
    double Dp = fma(2.0, u, -Gp), Dm = fma(2.0, v, -Gm);
    s.gap = phiP * fma(na, Gp, Hap);
    s.gam = phiM * fma(-na, Gm, Ham);
    s.gbp = phiP * fma(nb, Gp, Hbp);
    s.gbm = phiM * fma(-nb, Gm, Hbm);
    s.gcp = phiP * fma(sinvd, Dp, Hcp);
    s.gcm = phiM * fma(sinvd, Dm, Hcm);
    s.fp = phiP * Gp;
    s.fm = phiM * Gm;
}

template <int BLOCKS>
__global__ void __launch_bounds__(128, BLOCKS) k_chain(int steps, double seed, double* out) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    EState s{1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double sc = 1.0;
    const double wx = 0.7 + 1e-3 * (tid & 7), wy = 1.0, t = 0.3 + seed, na = 2.0, nb = 0.05, sinvd = 0.02;
    double r = 0.9 + 1e-4 * (tid & 31);
    for (int i = 0; i < steps; i += 2) {
        site(s, wx, wy, t, na, nb, sinvd, r, sc);
        sc = 1.0;
        site(s, wy, wx, t, na, nb, sinvd, r, sc);
        sc = inv_pow2_of(s.fp + s.fm);
        r = 1.8 - r;
    }
    if (s.fp + s.gap + s.gbm + s.gcp == 123.456) out[tid] = s.fm + s.gam + s.gbp + s.gcm;
}

// Variant with the real operand path: per step three 16-byte broadcast loads of the group's potentials and one 8-byte per-lane load of
// the emission weight, all from shared memory (filled once), 64-group tile walked round and round.
struct Tile { double2 w[64], tn[64], bd[64]; };
template <int BLOCKS>
__global__ void __launch_bounds__(128, BLOCKS) k_chain_smem(int steps, double seed, double* out) {
    __shared__ Tile tiles[4];
    __shared__ double ring[4][8][32];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = lane; j < 64; j += 32) {
        tiles[w].w[j] = make_double2(0.7 + 1e-3 * (j & 7), 1.0);
        tiles[w].tn[j] = make_double2(0.3 + seed, 2.0);
        tiles[w].bd[j] = make_double2(0.05, 0.02);
    }
    for (int u = 0; u < 8; ++u) ring[w][u][lane] = 0.9 + 1e-4 * lane + 0.2 * (u & 1);
    __syncwarp();
    EState s{1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double sc = 1.0;
    const Tile& st = tiles[w];
    for (int i = 0; i < steps; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = (i + u) & 63;
            const double2 ww = st.w[j], tn = st.tn[j], bd = st.bd[j];
            site(s, ww.x, ww.y, tn.x, tn.y, bd.x, bd.y, ring[w][u][lane], sc);
            if (u & 1) sc = inv_pow2_of(s.fp + s.fm); else sc = 1.0;
        }
    }
    if (s.fp + s.gap + s.gbm + s.gcp == 123.456) out[tid] = s.fm + s.gam + s.gbp + s.gcm;
}

// Variant V2: 40 instead of 48 bytes of warp-uniform operands per step — {e, t} and {nb, sinvd} as two 16-byte loads, N_l as one
// 8-byte load whose sign bit says which potential is the 1.0 (w = (e, 1) or (1, e)).
struct Tile2 { double2 et[64], bd[64]; double na[64]; };
template <int BLOCKS>
__global__ void __launch_bounds__(128, BLOCKS) k_chain_smem2(int steps, double seed, double* out) {
    __shared__ Tile2 tiles[4];
    __shared__ double ring[4][8][32];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = lane; j < 64; j += 32) {
        tiles[w].et[j] = make_double2(0.7 + 1e-3 * (j & 7), 0.3 + seed);
        tiles[w].bd[j] = make_double2(0.05, 0.02);
        tiles[w].na[j] = (j & 1) ? 2.0 : -2.0;
    }
    for (int u = 0; u < 8; ++u) ring[w][u][lane] = 0.9 + 1e-4 * lane + 0.2 * (u & 1);
    __syncwarp();
    EState s{1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double sc = 1.0;
    const Tile2& st = tiles[w];
    for (int i = 0; i < steps; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = (i + u) & 63;
            const double2 et = st.et[j], bd = st.bd[j];
            const double nas = st.na[j];
            const bool pos = __double2hiint(nas) >= 0;
            site(s, pos ? et.x : 1.0, pos ? 1.0 : et.x, et.y, fabs(nas), bd.x, bd.y, ring[w][u][lane], sc);
            if (u & 1) sc = inv_pow2_of(s.fp + s.fm); else sc = 1.0;
        }
    }
    if (s.fp + s.gap + s.gbm + s.gcp == 123.456) out[tid] = s.fm + s.gam + s.gbp + s.gcm;
}

template <int BLOCKS>
__global__ void __launch_bounds__(128, BLOCKS) k_dfma(int iters, double seed, double* out) {
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
    if (z == 123.456) out[tid] = z;
}

template <class F>
static float time_ms(F&& launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    return best;
}

template <int BLOCKS>
static void run(int sms, double* out) {
    const int grid = sms * BLOCKS, steps = 200000, iters = 400000;
    const double lanes = (double)grid * 128;
    float t_chain = time_ms([&] { k_chain<BLOCKS><<<grid, 128>>>(steps, 0.0, out); });
    float t_dfma = time_ms([&] { k_dfma<BLOCKS><<<grid, 128>>>(iters, 1.0, out); });
    float t_smem = time_ms([&] { k_chain_smem<BLOCKS><<<grid, 128>>>(steps, 0.0, out); });
    float t_smem2 = time_ms([&] { k_chain_smem2<BLOCKS><<<grid, 128>>>(steps, 0.0, out); });
    // FP64 instructions per lane: chain step = 27 (+1 add for the rescale every second step, 1 multiply dropped when sc == 1: 26.5)
    const double chain_inst = lanes * steps * 26.5, dfma_inst = lanes * (double)iters * 8;
    printf("{\"blocks_per_sm\": %d, \"warps_per_sm\": %d, \"chain_step_Ginst_per_s\": %.1f, \"dfma_Ginst_per_s\": %.1f, \"chain_over_dfma\": %.3f, "
           "\"dfma_TFLOPs\": %.2f, \"chain_step_smem_operands_Ginst_per_s\": %.1f, \"smem_over_dfma\": %.3f, \"smem40B_over_dfma\": %.3f}\n",
           BLOCKS, BLOCKS * 4, chain_inst / t_chain / 1e6, dfma_inst / t_dfma / 1e6, (chain_inst / t_chain) / (dfma_inst / t_dfma),
           2 * dfma_inst / t_dfma / 1e9, chain_inst / t_smem / 1e6, (chain_inst / t_smem) / (dfma_inst / t_dfma), (chain_inst / t_smem2) / (dfma_inst / t_dfma));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double* out;
    cudaMalloc(&out, 1 << 26);
    printf("# %s, %d SMs\n", p.name, p.multiProcessorCount);
    run<2>(p.multiProcessorCount, out);
    run<4>(p.multiProcessorCount, out);
    run<5>(p.multiProcessorCount, out);
    run<8>(p.multiProcessorCount, out);
    return 0;
}
