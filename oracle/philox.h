/*
 * Philox4x32-10 counter-based RNG (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
 * SC'11) -- ORACLE COPY, TEST INFRASTRUCTURE ONLY.  The product has its own device-side
 * implementation (swiftlink_b200/csrc/slk_philox.cuh); both are pinned to the Random123
 * known-answer vectors in tests/test_philox.py.
 *
 * Draw schedule shared by the oracle and the CUDA path (the reference uses a per-thread GSL
 * mt19937, random.cc:93-95, whose stream cannot be reproduced under a different partitioning;
 * north_star replaces it by a generator keyed by (chain, iteration, locus)):
 *
 *   key     = (seed_lo, seed_hi ^ chain)
 *   counter = (iteration_lo, iteration_hi, locus, slot >> 1)
 *   draw    = 53-bit uniform in [0,1) from words (2*(slot&1), 2*(slot&1)+1) of the block
 *
 * slot = peel-op index for the genotype draw of that op, nops + 2*person + parent for the
 * meiosis-indicator draw, and SLK_SLOT_PHASE for the per-sweep even/odd order draw (locus = 0).
 * M-sampler: ORC_SLOT_MEIOSIS + meiosis (= 2*(person - F) + parent) for the indicator draw of that
 * meiosis at `locus`; ORC_SLOT_MSHUFFLE with locus = i for step i of the Fisher-Yates shuffle of
 * the sweep's meiosis order; ORC_SLOT_KIND (locus = 0) for the L-sweep / M-sweep choice of an
 * iteration (markov_chain.cc:332).
 */
#ifndef SLK_ORACLE_PHILOX_H
#define SLK_ORACLE_PHILOX_H

#include <stdint.h>

#define ORC_PHILOX_M0 0xD2511F53u
#define ORC_PHILOX_M1 0xCD9E8D57u
#define ORC_PHILOX_W0 0x9E3779B9u
#define ORC_PHILOX_W1 0xBB67AE85u
#define ORC_SLOT_PHASE 0x7ffffff0u
#define ORC_SLOT_MSHUFFLE 0x7ffffff1u
#define ORC_SLOT_KIND 0x7ffffff2u
#define ORC_SLOT_MEIOSIS 0x40000000u

static inline void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for(int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t) ORC_PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t) ORC_PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t) p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t) p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += ORC_PHILOX_W0;
        k1 += ORC_PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline double orc_uniform(uint64_t seed, uint32_t chain, uint64_t iteration, uint32_t locus, uint32_t slot) {
    uint32_t key[2] = { (uint32_t) seed, (uint32_t)(seed >> 32) ^ chain };
    uint32_t ctr[4] = { (uint32_t) iteration, (uint32_t)(iteration >> 32), locus, slot >> 1 };
    uint32_t out[4];
    orc_philox4x32_10(ctr, key, out);
    uint32_t a = out[2 * (slot & 1)], b = out[2 * (slot & 1) + 1];
    uint64_t bits = (((uint64_t) a << 32) | b) >> 11;
    return (double) bits * (1.0 / 9007199254740992.0);
}

#endif
