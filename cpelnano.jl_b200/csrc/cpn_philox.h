// cpn_philox.h — counter-based generation of EM starts and read-label permutations.
//
// The reference draws both from Julia's global MersenneTwister: gen_ϕ0s (src/em_algorithm.jl:65-78) on every get_ϕhat! call
// (:374) and sort!(StatsBase.sample(1:n, n1; replace=false)) per permutation (src/two_samp_perm_tests.jl:211-214).  For the
// two-sample test that is 480 KB of starts and 120 KB of indices per region — pure RNG output, two orders of magnitude more
// than the region's data — so the library can regenerate them on the device from a 64-bit seed instead of receiving them.
// The stream is defined here (and restated independently by the CPU test oracle, so "bit-exact counts under a fixed seed" is
// well defined between oracle and GPU; parity with Julia's MersenneTwister stream is not attainable and not claimed):
//
//   Philox-4x32-10 (Salmon et al., SC'11), key = (seed low word, seed high word), counter = (c0, c1, c2, c3)
//   EM starts of unit (region r, permutation i, half h), start s:
//       counter = (r, i, h·65536 + s, 0)        i = 0xFFFFFFFF for the plain single-sample fit
//       (x0, x1, x2, ·) → k_a = ⌊x0·1001 / 2^32⌋, k_b = ⌊x1·10001 / 2^32⌋, k_c = ⌊x2·1001 / 2^32⌋
//       ϕ0 = ((k_a − 500)/100, (k_b − 5000)/100, k_c/100)      — the grid a ∈ −5:0.01:5, b ∈ −50:0.01:50, c ∈ 0:0.01:10;
//       integer-then-divide gives the correctly rounded decimals that Julia's StepRangeLen elements are
//   permutation i of region r, n pooled reads of which n1 form sample 1: selection sampling (Fan et al. 1962 / Knuth 3.4.2 S),
//       which visits the reads in order and therefore yields the ascending index set directly — the law of
//       sort!(sample(1:n, n1; replace=false)):
//       need = n1; for q = 0 … n−1: x = word (q mod 4) of the block with counter (r, i, q div 4, 1);
//                                   read q joins sample 1 iff ⌊x·(n − q) / 2^32⌋ < need (then need −= 1)
//   r is the 0-based position of the region in the call's batch.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define CPN_HD __host__ __device__ __forceinline__
#else
#define CPN_HD inline
#endif

struct CpnPhilox4 { uint32_t x[4]; };

CPN_HD uint32_t cpn_mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }

CPN_HD CpnPhilox4 cpn_philox4x32_10(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    CpnPhilox4 r;
    r.x[0] = c0; r.x[1] = c1; r.x[2] = c2; r.x[3] = c3;
    return r;
}

constexpr uint32_t CPN_RNG_NO_PERM = 0xFFFFFFFFu;

// EM start s of unit (region, perm, half): gen_ϕ0s' grid (src/em_algorithm.jl:72)
CPN_HD void cpn_gen_phi0(uint64_t seed, uint32_t region, uint32_t perm, uint32_t half, uint32_t start, double out[3]) {
    const CpnPhilox4 r = cpn_philox4x32_10(seed, region, perm, half * 65536u + start, 0u);
    out[0] = (double)((int)cpn_mulhi32(r.x[0], 1001u) - 500) / 100.0;
    out[1] = (double)((int)cpn_mulhi32(r.x[1], 10001u) - 5000) / 100.0;
    out[2] = (double)cpn_mulhi32(r.x[2], 1001u) / 100.0;
}
