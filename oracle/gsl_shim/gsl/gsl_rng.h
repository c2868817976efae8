/*
 * Minimal stand-in for <gsl/gsl_rng.h> -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference (ajm/swiftlink) uses GSL solely for an mt19937 stream in src/random.cc
 * (gsl_rng_env_setup, gsl_rng_default, gsl_rng_alloc/free/set/name, gsl_rng_uniform,
 * gsl_rng_uniform_int).  GSL is not installed in this image, so the oracle build
 * (oracle/Makefile) puts this directory on the include path when compiling the reference
 * sources where they lie under /root/reference/src.  Semantics follow GSL's documented
 * behaviour for mt19937: seed 0 is replaced by 4357, uniform() = next()/2^32 in [0,1),
 * uniform_int(n) = rejection sampling on next()/(0xffffffff/n).
 *
 * Nothing here is linked into the product library.
 */
#ifndef SLK_ORACLE_GSL_RNG_SHIM_H
#define SLK_ORACLE_GSL_RNG_SHIM_H

#include <stdlib.h>

typedef struct { const char* name; } gsl_rng_type;

typedef struct {
    const gsl_rng_type* type;
    unsigned long mt[624];
    int mti;
} gsl_rng;

static const gsl_rng_type slk_shim_mt19937_type = { "mt19937" };
static const gsl_rng_type* gsl_rng_default = &slk_shim_mt19937_type;

static inline const gsl_rng_type* gsl_rng_env_setup(void) {
    gsl_rng_default = &slk_shim_mt19937_type;
    return gsl_rng_default;
}

static inline void gsl_rng_set(gsl_rng* r, unsigned long seed) {
    if(seed == 0) seed = 4357;
    r->mt[0] = seed & 0xffffffffUL;
    for(int i = 1; i < 624; ++i) {
        r->mt[i] = (1812433253UL * (r->mt[i-1] ^ (r->mt[i-1] >> 30)) + (unsigned long)i) & 0xffffffffUL;
    }
    r->mti = 624;
}

static inline gsl_rng* gsl_rng_alloc(const gsl_rng_type* t) {
    gsl_rng* r = (gsl_rng*) malloc(sizeof(gsl_rng));
    r->type = t;
    gsl_rng_set(r, 0);
    return r;
}

static inline void gsl_rng_free(gsl_rng* r) { free(r); }
static inline const char* gsl_rng_name(const gsl_rng* r) { return r->type->name; }

static inline unsigned long slk_shim_mt_next(gsl_rng* r) {
    const unsigned long UPPER = 0x80000000UL, LOWER = 0x7fffffffUL;
    unsigned long* mt = r->mt;
    if(r->mti >= 624) {
        int kk;
        for(kk = 0; kk < 624 - 397; ++kk) {
            unsigned long y = (mt[kk] & UPPER) | (mt[kk+1] & LOWER);
            mt[kk] = mt[kk+397] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
        }
        for(; kk < 623; ++kk) {
            unsigned long y = (mt[kk] & UPPER) | (mt[kk+1] & LOWER);
            mt[kk] = mt[kk+(397-624)] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
        }
        {
            unsigned long y = (mt[623] & UPPER) | (mt[0] & LOWER);
            mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
        }
        r->mti = 0;
    }
    unsigned long k = mt[r->mti++];
    k ^= (k >> 11);
    k ^= (k << 7) & 0x9d2c5680UL;
    k ^= (k << 15) & 0xefc60000UL;
    k ^= (k >> 18);
    return k & 0xffffffffUL;
}

static inline double gsl_rng_uniform(gsl_rng* r) {
    return slk_shim_mt_next(r) / 4294967296.0;
}

static inline unsigned long gsl_rng_uniform_int(gsl_rng* r, unsigned long n) {
    unsigned long scale = 0xffffffffUL / n;
    unsigned long k;
    do { k = slk_shim_mt_next(r) / scale; } while(k >= n);
    return k;
}

#endif
