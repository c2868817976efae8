// slk_philox.cuh -- Philox4x32-10 (Salmon et al., SC'11) for the L-sampler draws.
//
// The reference draws from a per-OpenMP-thread GSL mt19937 (random.cc:93-95), and its stale GPU
// sampler from one XORWOW state per block (cuda_random.cu:7-22): both make the sampled chain
// depend on how work is partitioned.  north_star replaces them with a counter-based generator
// keyed by (chain, iteration, locus):
//
//   key     = (seed_lo, seed_hi ^ chain)
//   counter = (iteration_lo, iteration_hi, locus, slot >> 1)
//   draw    = 53-bit uniform in [0,1) from words (2*(slot&1), 2*(slot&1)+1)
//
// slot = peel-op index for that op's genotype draw, nops + 2*person + parent for a meiosis
// indicator, SLK_SLOT_PHASE (with locus 0) for the per-sweep parity-order draw.
#ifndef SLK_PHILOX_CUH
#define SLK_PHILOX_CUH

#include <stdint.h>

#define SLK_SLOT_PHASE 0x7ffffff0u

// Partition of the `iteration` counter.  MCMC iterations count up from 0; the sequential-imputation runs that build
// a chain's start graph (and the locus-by-locus fallback) are numbered from SLK_SI_FIRST_RUN, so the two never share
// a (seed, chain, iteration, locus, slot) tuple -- with a common numbering a chain started with `-b 0` would redraw,
// in its first sweeps, the uniforms that had produced its start graph.
#define SLK_SI_FIRST_RUN (1ull << 62)

#if defined(__CUDACC__)
#define SLK_HD __host__ __device__ __forceinline__
#else
#define SLK_HD static inline
#endif

SLK_HD void slk_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                           uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for(int r = 0; r < 10; ++r) {
#if defined(__CUDA_ARCH__)
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
        uint64_t p0 = (uint64_t) M0 * c0, p1 = (uint64_t) M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t) p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t) p1;
#endif
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

SLK_HD double slk_uniform(uint64_t seed, uint32_t chain, uint64_t iteration,
                                                       uint32_t locus, uint32_t slot) {
    uint32_t o[4];
    slk_philox4x32_10((uint32_t) iteration, (uint32_t)(iteration >> 32), locus, slot >> 1,
                      (uint32_t) seed, (uint32_t)(seed >> 32) ^ chain, o);
    uint32_t a = (slot & 1) ? o[2] : o[0];
    uint32_t b = (slot & 1) ? o[3] : o[1];
    uint64_t bits = (((uint64_t) a << 32) | b) >> 11;
    return (double) bits * (1.0 / 9007199254740992.0);
}

#endif
