// slk_peel.h -- the per-thread work of the peel kernels: one forward tile, one backward term, the
// per-locus staging of one op / one meiosis / one indicator.  Plain C++ over pointers, so the very
// same code is compiled into the sm_100a kernels (slk_kernels.cuh) and into the sequential CPU
// emulation the tests use to check the plan flattening and the index arithmetic without a GPU
// (tests/emu/slk_emu.cc -- test infrastructure, never linked into the product library).
//
// A forward TILE is one row (one valid assignment of cutset digits 1..c-1) x the four values of
// cutset digit 0 (slots) x the four genotypes of the peel node (variants): 16 product chains in
// registers.  Matrices are indexed by the SORTED cutset (slk_types.h), so a consumed function is
// read as 16, 4 or 1 consecutive doubles (SLK_KEY_*) and the result is written as 4 consecutive
// doubles.
//
// Arithmetic follows the reference's operation order (compiled with -fmad=false / -ffp-contract=off)
// so peel matrices are bit-identical to the CPU R-functions: rfunction.cc:116-142,
// sampler_rfunction.cc:102-157,190-287,350-419, trait_rfunction.cc:9-137.
#ifndef SLK_PEEL_H
#define SLK_PEEL_H

#include <stdint.h>

#include "swiftlink_b200.h"
#include "slk_types.h"

#if defined(__CUDACC__)
#define SLK_DEV __device__ __forceinline__
#else
#define SLK_DEV static inline
#endif

// what a unit of work needs besides the program (filled by the kernel / the emulation)
struct SlkEnv {
    int N, F, nops;
    int sex_linked;
    const uint8_t* male;         // [N]
    // parity dumps (DBG instantiations only)
    double* dump_mat;            // dense matrices, reference cell order (pre-zeroed by the host)
    double* dump_pre;            // dense presum matrices
    const uint32_t* dbg;         // [nops][3] dense offset, refpos
};

struct SlkProgView {             // CTA-shared (shared memory on the device)
    const uint32_t* stream;
    const uint16_t* op_start;
    const uint16_t* imap;
    const uint32_t* flevel_items;
    const uint32_t* flevel_map;
    const uint16_t* bops;
    const uint16_t* blevel_start;
    const uint8_t* glist;        // [16] ascending list of the genotypes present in a 4-bit mask, 2 bits each
    const double* dprob;         // [N][4] disease probabilities
};

struct SlkTeam {                 // per team
    double* arena;               // shared part of the arena
    double* garena;              // global part (arena offsets >= smem_doubles)
    int smem_doubles;
    double* tables;              // (N-F) x 8 transmission tables (sampler) / (N-F) x 4 recombination weights (trait)
    double* scal;                // 48: thetas [0..4), class priors [16..36), founder priors of the locus [36..44)
    SlkOpLoc* oploc;             // [nops]
    double* ru;                  // [nops] genotype draws of the current locus (sampler)
    uint8_t* gc;                 // gcode row of this locus
    uint8_t* dgl;                // descent graph at the left neighbour / interval start
    uint8_t* dgr;                // right neighbour / interval end
    uint8_t* pmk;                // sampled genotypes
    int* red;
};

static const uint8_t kSlkGlist[16] = {
    0x00, 0x00, 0x01, 0x04, 0x02, 0x08, 0x09, 0x24, 0x03, 0x0C, 0x0D, 0x34, 0x0E, 0x38, 0x39, 0xE4
};

#if defined(__CUDACC__)
static __constant__ double kSlkClassPrior[5][4] = {
#else
static const double kSlkClassPrior[5][4] = {
#endif
    {1.0, 0.0, 0.0, 0.0},
    {0.0, 1.0, 0.0, 0.0},
    {0.0, 0.0, 0.5, 0.5},
    {0.25, 0.25, 0.25, 0.25},
    {0.5, 0.5, 0.0, 0.0}
};

// ---- small helpers ---------------------------------------------------------------------------

SLK_DEV uint32_t slk_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t) __popc(x);
#else
    return (uint32_t) __builtin_popcount(x);
#endif
}

SLK_DEV uint32_t slk_div3(uint32_t j) {
#if defined(__CUDA_ARCH__)
    return __umulhi(j, 0x55555556u);
#else
    return j / 3u;
#endif
}

// four consecutive doubles (the address is 16-byte aligned by construction)
SLK_DEV void slk_ld4(const double* p, double x[4]) {
#if defined(__CUDA_ARCH__)
    const double2 a = *reinterpret_cast<const double2*>(p);
    const double2 b = *reinterpret_cast<const double2*>(p + 2);
    x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
#else
    x[0] = p[0]; x[1] = p[1]; x[2] = p[2]; x[3] = p[3];
#endif
}

SLK_DEV void slk_st4(double* p, const double x[4]) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<double2*>(p) = make_double2(x[0], x[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(x[2], x[3]);
#else
    p[0] = x[0]; p[1] = x[1]; p[2] = x[2]; p[3] = x[3];
#endif
}

SLK_DEV double slk_sel4(const double v[4], uint32_t i) {
    const double a = (i & 1u) ? v[1] : v[0];
    const double b = (i & 1u) ? v[3] : v[2];
    return (i & 2u) ? b : a;
}

// element offset inside a matrix: the padded layout inserts two doubles after every 16
SLK_DEV uint32_t slk_pidx(uint32_t idx, bool pad) { return pad ? idx + ((idx >> 4) << 1) : idx; }

SLK_DEV double* slk_mat_ptr(const SlkTeam& tm, uint32_t offw) {
    const int off = (int)(offw & SLK_OFF_MASK);
    return off < tm.smem_doubles ? tm.arena + off : tm.garena + (off - tm.smem_doubles);
}

SLK_DEV uint32_t slk_nvalid(unsigned long long lmw, int first, int c) {
    uint32_t n = 1;
    for(int d = first; d < c; ++d) n *= slk_popc((uint32_t)(lmw >> (4 * d)) & 15u);
    return n;
}

// row -> the packed digits 1..c-1 of the j-th valid row (digit d at bits [2d, 2d+2)): a mixed-radix decode of j
// over the per-digit legal sets
SLK_DEV uint32_t slk_decode_row(uint32_t j, unsigned long long lmw, int c, const uint8_t* glist) {
    uint32_t cell = 0;
    for(int d = 1; d < c; ++d) {
        const uint32_t lm = (uint32_t)(lmw >> (4 * d)) & 15u;
        const uint32_t k = slk_popc(lm);
        const uint32_t q = (k == 3u) ? slk_div3(j) : (j >> (k >> 1));       // k = 1, 2, 4: shift by 0, 1, 2
        const uint32_t r = j - q * k;
        cell |= (((uint32_t) glist[lm] >> (2u * r)) & 3u) << (2 * d);
        j = q;
    }
    return cell;
}

SLK_DEV unsigned long long slk_full_lmw(int c) {
    return (c >= 16) ? ~0ull : ((1ull << (4 * c)) - 1ull);
}

// row index of a consumed matrix: up to five runs of consecutive digits of the consumer's cell
SLK_DEV uint32_t slk_run(uint32_t cell, uint32_t run) {
    return ((cell >> (run & 31u)) & ((1u << ((run >> 10) & 31u)) - 1u)) << ((run >> 5) & 31u);
}

SLK_DEV uint32_t slk_gather(uint32_t cell, const uint32_t* pw) {
    const uint32_t w1 = pw[1], w2 = pw[2];
    const uint32_t nruns = (w1 >> 8) & 0xffu;
    uint32_t R = slk_run(cell, w2);                     // an absent run is all zeros: no bits
    if(nruns > 1u) {
        R |= slk_run(cell, w2 >> 16);
        if(nruns > 2u) {
            const uint32_t w3 = pw[3];
            R |= slk_run(cell, w3) | slk_run(cell, w3 >> 16) | slk_run(cell, w1 >> 16);
        }
    }
    return R;
}

// Per-child transmission tables (sampler_rfunction.cc:102-157), staged at the start of the locus: for each parent
// (0 maternal, 1 paternal) the four entries that depend on the neighbouring meioses,
//     tb[4 p + 0..1] = P(allele 0 / 1 | parent AU),  tb[4 p + 2..3] = P(allele 0 / 1 | parent UA);
// a homozygous parent transmits its allele with certainty.  Under X linkage a heterozygous father transmits
// nothing (entry 0) and a son takes nothing from his father (both paternal entries of a homozygous father are 1).
SLK_DEV double slk_tb_entry(const double* tb, uint32_t parent, uint32_t pg, uint32_t a, bool xpat, bool xson) {
    const double het = tb[4u * parent + 2u * (pg & 1u) + a];
    const double hom = (a == pg || xson) ? 1.0 : 0.0;
    return pg >= 2u ? (xpat ? 0.0 : het) : hom;
}

// sampler_rfunction.cc:350-419: one entry of the 64-entry transmission table = maternal entry x paternal entry;
// a son is never heterozygous
SLK_DEV double slk_trans_prob(bool sex_linked, bool male, uint32_t m, uint32_t f, uint32_t kg, const double* tb) {
    const uint32_t am = ((kg >> 1) ^ kg) & 1u;       // maternal allele of the kid: AA, AU carry A
    const uint32_t ap = kg & 1u;                     // paternal allele: AA, UA carry A
    if(sex_linked && male && kg >= 2u) return 0.0;
    return slk_tb_entry(tb, 0u, m, am, false, false) * slk_tb_entry(tb, 1u, f, ap, sex_linked, sex_linked && male);
}

// rfunction.h:81-107
SLK_DEV bool slk_affected_trait(uint32_t pt, uint32_t allele) {
    if(pt == SLK_UU) return false;
    if(pt == SLK_AA) return true;
    return (pt == SLK_AU) ? (allele == 0) : (allele == 1);
}

// rfunction.cc:94-113
SLK_DEV uint32_t slk_phased_trait(bool sex_linked, bool male, uint32_t m, uint32_t f, uint32_t i, uint32_t j) {
    const bool ma = slk_affected_trait(m, i), pa = slk_affected_trait(f, j);
    if(sex_linked && male) return ma ? SLK_AA : SLK_UU;
    if(ma) return pa ? SLK_AA : SLK_AU;
    return pa ? SLK_UA : SLK_UU;
}

// the four values of slk_phased_trait(m, f, i, j), (i, j) = (0,0), (0,1), (1,0), (1,1), packed two bits each: one
// byte per (m, f), looked up in four 32-bit constants (byte f of word m) generated from the function above
SLK_DEV uint32_t slk_phased_trait4(bool sex_linked, bool male, uint32_t m, uint32_t f) {
    uint32_t w;
    if(sex_linked && male) w = (m == 0u) ? 0x00000000u : (m == 1u) ? 0x55555555u : (m == 2u) ? 0x05050505u : 0x50505050u;
    else                   w = (m == 0u) ? 0xcc33ff00u : (m == 1u) ? 0x669955aau : (m == 2u) ? 0xc639f50au : 0x6c935fa0u;
    return (w >> (8u * f)) & 0xffu;
}

// trait_rfunction.cc:106-127: sum, in (i, j) order, of the recombination weights of the
// transmitted-allele choices that produce the kid's genotype.  The choices factorise: the kid's
// genotype fixes whether the maternal (paternal) allele must be the A allele, and the parent's
// genotype fixes which of its two alleles (i = 0, 1) qualify -- a 2-bit mask per parent, read
// from a packed table indexed by (parent genotype, required) [rfunction.h:81-107].
SLK_DEV double slk_trait_child_sum(bool sex_linked, bool male, uint32_t m, uint32_t f, uint32_t kg, const double* w) {
    const uint32_t LUT = 0x96C3u;       // nibble per parent genotype: (mask if A required) << 2 | (mask if U required)
    uint32_t mi, mj;
    if(sex_linked && male) {
        if(kg >= 2u) return 0.0;                       // a son is never heterozygous (rfunction.cc:101-103)
        mi = (LUT >> (4u * m + 2u * (kg == SLK_AA ? 1u : 0u))) & 3u;
        mj = 3u;
    }
    else {
        const uint32_t ma = ((kg >> 1) ^ kg) & 1u;     // AA, AU: maternal allele is A
        const uint32_t pa = kg & 1u;                   // AA, UA: paternal allele is A
        mi = (LUT >> (4u * m + 2u * ma)) & 3u;
        mj = (LUT >> (4u * f + 2u * pa)) & 3u;
    }
    double s = 0.0;
    if((mi & 1u) && (mj & 1u)) s += w[0];
    if((mi & 1u) && (mj & 2u)) s += w[1];
    if((mi & 2u) && (mj & 1u)) s += w[2];
    if((mi & 2u) && (mj & 2u)) s += w[3];
    return s;
}

// prior of the peel node's four genotypes: marker prior classes at the staged locus (person.cc:224-245), or the
// disease probabilities (trait program; ELOD's simulated trait locus, person.h:204-208)
template<bool TRAIT>
SLK_DEV void slk_load_prior(const SlkProgView& pg, const SlkTeam& tm, int person, double tc[4]) {
    if(TRAIT) { slk_ld4(pg.dprob + 4 * person, tc); return; }
    const uint32_t cls = tm.gc[person] >> 4;
    if(cls == SLK_PRIOR_PERSON) {
#pragma unroll
        for(int g = 0; g < 4; ++g) tc[g] = pg.dprob[4 * person + g];
        return;
    }
    slk_ld4(tm.scal + 16 + cls * 4, tc);               // classes 0-4 fixed, 5-6 founder priors of this locus
}

// ---- per-locus staging of one op: legal masks of the sorted cutset, valid rows --------------------
template<bool TRAIT>
SLK_DEV void slk_stage_oploc(const SlkProgView& pg, const SlkTeam& tm, int op) {
    const uint32_t* rec = pg.stream + 4u * pg.op_start[op];
    const int c = (rec[0] >> 4) & 15u;
    const int peel = rec[0] >> 16;
    const uint32_t* cw = rec + SLK_REC_HEADER;
    unsigned long long w = 0;
    uint32_t peel_lm;
    if(TRAIT) {
        const uint32_t* lw = cw + ((c + 1) >> 1);
        w = (unsigned long long) lw[0] | ((unsigned long long) lw[1] << 32);
        peel_lm = (rec[3] >> 8) & 15u;
    }
    else {
        for(int d = 0; d < c; ++d) {
            const uint32_t person = (cw[d >> 1] >> (16 * (d & 1))) & 0xffffu;
            w |= (unsigned long long)(tm.gc[person] & 15u) << (4 * d);
        }
        peel_lm = (uint32_t) tm.gc[peel] & 15u;
    }
    SlkOpLoc ol;
    ol.lmw = w;
    ol.nrows = slk_nvalid(w, 1, c);
    ol.misc = (c > 0 ? ((uint32_t) w & 15u) : 1u) | (peel_lm << 4);
    tm.oploc[op] = ol;
}

// ---- one forward tile ------------------------------------------------------------------------------
//
// TRAIT = false: SamplerRfunction::evaluate_element, TRAIT = true: TraitRfunction::evaluate_element
// DBG adds the dense parity dumps (separate instantiation, keeps the hot kernels lean).
// NS = 4: the thread owns the whole tile (throughput mode, big levels); NS = 1: the thread owns slot s0 only
// (latency mode: a level with few rows is spread over four times as many threads)
template<bool TRAIT, bool DBG, int NS>
SLK_DEV void slk_forward_tile(const SlkEnv& E, const SlkProgView& pg, const SlkTeam& tm, const uint32_t* rec, uint32_t row,
                              uint32_t s0) {
    const uint32_t w0 = rec[0], matw = rec[1], w3 = rec[3];
    const uint32_t op = w3 >> 16;
    const SlkOpLoc ol = tm.oploc[op];
    if(row >= ol.nrows) return;
    const bool sex_linked = E.sex_linked != 0;
    const int type = w0 & 7u, c = (w0 >> 4) & 15u, nprev = (w0 >> 8) & 15u, nkids = (w0 >> 12) & 15u;
    const int peel = w0 >> 16;
    const uint32_t lm0 = ol.misc & 15u, peel_lm = (ol.misc >> 4) & 15u;
    const unsigned long long lmw_hi = ol.lmw >> 4;
    const int ch = c > 0 ? c - 1 : 0;
    // digits 1..c-1 of this row, digit d at bits [2d, 2d+2); digit 0 is the slot axis
    const uint32_t cell = (lmw_hi == slk_full_lmw(ch)) ? (row << 2) : slk_decode_row(row, ol.lmw, c, pg.glist);
    const uint32_t* pw = rec + SLK_REC_HEADER + ((c + 1) >> 1) + (TRAIT ? 2 : 0);
    const uint32_t* kw = pw + SLK_REC_PREV * nprev;

    double tc[4];
    slk_load_prior<TRAIT>(pg, tm, peel, tc);
    if(!TRAIT && (w3 & 1u)) {
        // a previous function keyed on the peel node is 0 wherever that genotype was eliminated (its cell is
        // not in valid_indices): the product is 0
#pragma unroll
        for(int g = 0; g < 4; ++g) if(!((peel_lm >> g) & 1u)) tc[g] = 0.0;
    }

    // digit value of sorted cutset position `pos` in slot s of this thread
#define SLK_DIG(pos, s) (((pos) == 0u) ? (s0 + (uint32_t)(s)) : ((cell >> (2u * (pos))) & 3u))

    // t[s][v]: the product chain of slot s and peel-node GENOTYPE v.  In a trait child peel the reference's four chains
    // are the (maternal, paternal) transmitted-allele choices (trait_rfunction.cc:44-66); a choice's chain depends on
    // the choice only through the kid's genotype it implies, factor by factor, so the four genotype chains are run
    // instead (same factors, same order: bit-identical values) and the choices pick theirs at the end.
    double t[NS][4];
    // ---- initial factor: prior of the peel-node genotype (x transmission for a sampler child peel)
    const uint32_t k0 = (type == SLK_CHILD_PEEL) ? kw[0] : 0u;
    const uint32_t mpos0 = (k0 >> 20) & 15u, fpos0 = (k0 >> 24) & 15u;
    const bool male0 = (k0 >> 28) & 1u;
#pragma unroll
    for(int s = 0; s < NS; ++s) {
#pragma unroll
        for(int v = 0; v < 4; ++v) t[s][v] = tc[v];
    }
    if(!TRAIT && type == SLK_CHILD_PEEL) {
        // T[m][f][v] = P(maternal allele of v | m) * P(paternal allele of v | f) (sampler_rfunction.cc:412-415); the
        // four factors are the same in every slot unless a parent is cutset digit 0
        const double* tb = tm.tables + 8 * (peel - E.F);
        const bool xson = sex_linked && male0;
        const bool dep_s = NS > 1 && (mpos0 == 0u || fpos0 == 0u);
        double T[4];
#pragma unroll
        for(int s = 0; s < NS; ++s) {
            if(s == 0 || dep_s) {
                const uint32_t m = SLK_DIG(mpos0, s), f = SLK_DIG(fpos0, s);
                const double m0 = slk_tb_entry(tb, 0u, m, 0u, false, false), m1 = slk_tb_entry(tb, 0u, m, 1u, false, false);
                const double f0 = slk_tb_entry(tb, 1u, f, 0u, sex_linked, xson), f1 = slk_tb_entry(tb, 1u, f, 1u, sex_linked, xson);
                T[SLK_UU] = m0 * f0;
                T[SLK_AA] = m1 * f1;
                T[SLK_AU] = xson ? 0.0 : m1 * f0;          // a son is never heterozygous
                T[SLK_UA] = xson ? 0.0 : m0 * f1;
            }
#pragma unroll
            for(int v = 0; v < 4; ++v) t[s][v] *= T[v];
        }
    }

    // ---- previous functions, in the reference's order
    for(int jp = 0; jp < nprev; ++jp, pw += SLK_REC_PREV) {
        const uint32_t offw0 = pw[0];
        const uint32_t kind = pw[1] & 0xffu;
        const bool pad0 = (offw0 & SLK_MAT_PAD) != 0u;
        const uint32_t R0 = slk_gather(cell, pw);
        const double* p = slk_mat_ptr(tm, offw0) + (kind == SLK_KEY_VS ? (pad0 ? 18u * R0 : 16u * R0) + 4u * s0
                                                    : kind == SLK_KEY_R ? slk_pidx(R0, pad0) : slk_pidx(4u * R0, pad0));
        if(kind == SLK_KEY_VS) {
#pragma unroll
            for(int s = 0; s < NS; ++s) {
                double x[4];
                slk_ld4(p + 4 * s, x);
#pragma unroll
                for(int v = 0; v < 4; ++v) t[s][v] *= x[v];
            }
        }
        else if(kind == SLK_KEY_V) {
            double x[4];
            slk_ld4(p, x);
#pragma unroll
            for(int s = 0; s < NS; ++s)
#pragma unroll
                for(int v = 0; v < 4; ++v) t[s][v] *= x[v];
        }
        else if(kind == SLK_KEY_S) {
            double x[4];
            if(NS == 4) slk_ld4(p, x);
            else x[0] = p[s0];
#pragma unroll
            for(int s = 0; s < NS; ++s)
#pragma unroll
                for(int v = 0; v < 4; ++v) t[s][v] *= x[s];
        }
        else {
            const double x = p[0];
#pragma unroll
            for(int s = 0; s < NS; ++s)
#pragma unroll
                for(int v = 0; v < 4; ++v) t[s][v] *= x;
        }
    }

    // ---- transmission to the children / recombination weights
    if(TRAIT && type == SLK_CHILD_PEEL) {
        // choice (i, j) takes the chain of the genotype it implies, times its recombination weight
        double w[4];
        slk_ld4(tm.tables + 4 * (peel - E.F), w);
#pragma unroll
        for(int s = 0; s < NS; ++s) {
            const uint32_t g4 = slk_phased_trait4(sex_linked, male0, SLK_DIG(mpos0, s), SLK_DIG(fpos0, s));
            double u[4];
#pragma unroll
            for(int v = 0; v < 4; ++v) u[v] = slk_sel4(t[s], (g4 >> (2 * v)) & 3u) * w[v];
#pragma unroll
            for(int v = 0; v < 4; ++v) t[s][v] = u[v];
        }
    }
    else if(type == SLK_PARENT_PEEL) {
        // child_prob = prod_k T_k (sampler_rfunction.cc:258-279 / trait_rfunction.cc:91-129).  When no child's own or
        // other parent's digit is cutset digit 0, the four products (one per peel genotype) are the same in every slot
        // and are evaluated once per tile.
        if(nkids > 0) {
            bool dep_s = false;
            if(NS > 1) {
                for(int k = 0; k < nkids; ++k) {
                    const uint32_t kd = kw[k];
                    const uint32_t mp = (kd >> 20) & 15u, fp = (kd >> 24) & 15u, kp = (kd >> 16) & 15u;
                    dep_s = dep_s || kp == 0u || (mp == SLK_POS_PEEL ? fp : mp) == 0u;
                }
            }
            double cp[4];
#pragma unroll
            for(int s = 0; s < NS; ++s) {
                if(s == 0 || dep_s) {
                    for(int k = 0; k < nkids; ++k) {
                        const uint32_t kd = kw[k];
                        const uint32_t mp = (kd >> 20) & 15u, fp = (kd >> 24) & 15u, kp = (kd >> 16) & 15u;
                        const bool male = (kd >> 28) & 1u;
                        const double* tb = tm.tables + (TRAIT ? 4 : 8) * ((int)(kd & 0xffffu) - E.F);
                        const bool peel_is_mother = mp == SLK_POS_PEEL;
                        const uint32_t kg = SLK_DIG(kp, s), og = SLK_DIG(peel_is_mother ? fp : mp, s);   // the other parent is a cutset member
                        double T[4];
                        if(TRAIT) {
#pragma unroll
                            for(int v = 0; v < 4; ++v)
                                T[v] = slk_trait_child_sum(sex_linked, male, peel_is_mother ? (uint32_t) v : og,
                                                           peel_is_mother ? og : (uint32_t) v, kg, tb);
                        }
                        else {
                            const uint32_t am = ((kg >> 1) ^ kg) & 1u, ap = kg & 1u;
                            const bool xson = sex_linked && male;
                            if(xson && kg >= 2u) { T[0] = T[1] = T[2] = T[3] = 0.0; }
                            else if(peel_is_mother) {
                                const double B = slk_tb_entry(tb, 1u, og, ap, sex_linked, xson);
#pragma unroll
                                for(int v = 0; v < 4; ++v) T[v] = slk_tb_entry(tb, 0u, (uint32_t) v, am, false, false) * B;
                            }
                            else {
                                const double B = slk_tb_entry(tb, 0u, og, am, false, false);
#pragma unroll
                                for(int v = 0; v < 4; ++v) T[v] = B * slk_tb_entry(tb, 1u, (uint32_t) v, ap, sex_linked, xson);
                            }
                        }
#pragma unroll
                        for(int v = 0; v < 4; ++v) cp[v] = (k == 0) ? T[v] : cp[v] * T[v];   // child_prob starts at 1.0: 1.0 * T == T
                    }
                }
#pragma unroll
                for(int v = 0; v < 4; ++v) t[s][v] *= cp[v];
            }
        }
    }

    // ---- the sum over the peel node; slots whose digit-0 genotype is illegal stay 0 as in the reference
    double out[NS];
#pragma unroll
    for(int s = 0; s < NS; ++s) {
        const double total = ((t[s][0] + t[s][1]) + t[s][2]) + t[s][3];
        out[s] = ((lm0 >> (s0 + s)) & 1u) ? total : 0.0;
    }
    double* Y = slk_mat_ptr(tm, matw);
    const bool ypad = (matw & SLK_MAT_PAD) != 0u;
    if(NS == 4) {
        if(c == 0) Y[0] = out[0];
        else slk_st4(Y + slk_pidx(cell, ypad), out);
    }
    else if(c > 0 || s0 == 0u) Y[slk_pidx(cell | s0, ypad)] = out[0];

    if(DBG) {
        const uint32_t* dbg = E.dbg + 3u * op;
        const int dense_off = (int) dbg[0];
        const unsigned long long refpos = (unsigned long long) dbg[1] | ((unsigned long long) dbg[2] << 32);
        for(int s = 0; s < NS; ++s) {
            if(!((lm0 >> (s0 + s)) & 1u)) continue;
            // the reference's cell index: sorted digit d sits at position refpos[d] of its cutset
            uint32_t ref = 0;
            for(int d = 0; d < c; ++d) ref |= SLK_DIG((uint32_t) d, s) << (2u * (uint32_t)((refpos >> (4 * d)) & 15u));
            if(E.dump_mat) E.dump_mat[dense_off + ref] = out[s];
            if(!TRAIT && E.dump_pre) {
#pragma unroll
                for(int v = 0; v < 4; ++v) E.dump_pre[4 * (size_t) dense_off + ref + ((size_t) v << (2 * c))] = t[s][v];
            }
        }
    }
#undef SLK_DIG
}

// one item of a forward level: finds the op by the item map, runs the tile (whole, or one slot in a fine level)
template<bool TRAIT, bool DBG>
SLK_DEV void slk_forward_item(const SlkEnv& E, const SlkProgView& pg, const SlkTeam& tm, const uint16_t* map, bool fine, uint32_t q) {
    const uint32_t ritem = fine ? (q >> 2) : q;
    const uint32_t* rec = pg.stream + 4u * map[ritem >> 2];
    const uint32_t row = ritem - rec[2];
    if(fine) slk_forward_tile<TRAIT, DBG, 1>(E, pg, tm, rec, row, q & 3u);
    else     slk_forward_tile<TRAIT, DBG, 4>(E, pg, tm, rec, row, 0u);
}

// the cell of a previous function at the sampled assignment `cell` of the consumer's cutset and genotype g of its peel node
SLK_DEV const double* slk_backward_cell(const SlkTeam& tm, const uint32_t* pw, uint32_t cell, uint32_t s, uint32_t g) {
    const uint32_t offw = pw[0];
    const uint32_t kind = pw[1] & 0xffu;
    const bool pad = (offw & SLK_MAT_PAD) != 0u;
    const uint32_t R = slk_gather(cell & ~3u, pw);
    uint32_t idx;
    if(kind == SLK_KEY_VS)     idx = g + 4u * s + 16u * R;
    else if(kind == SLK_KEY_V) idx = g + 4u * R;
    else if(kind == SLK_KEY_S) idx = s + 4u * R;
    else                       idx = R;
    return slk_mat_ptr(tm, offw) + slk_pidx(idx, pad);
}

// ---- one term of the backward pass: presum(cell, g) of op recomputed (sampler only) --------
// The presum matrix (rfunction.cc:20-21) is never stored; SamplerRfunction::sample (sampler_rfunction.cc:159-188)
// reads it at the already sampled cutset assignment only, four values per op.
SLK_DEV double slk_backward_term(const SlkEnv& E, const SlkProgView& pg, const SlkTeam& tm, uint32_t op, uint32_t g) {
    const bool sex_linked = E.sex_linked != 0;
    const uint32_t* rec = pg.stream + 4u * pg.op_start[op];
    const uint32_t w0 = rec[0];
    const int type = w0 & 7u, c = (w0 >> 4) & 15u, nprev = (w0 >> 8) & 15u, nkids = (w0 >> 12) & 15u;
    const int peel = w0 >> 16;
    const uint32_t* cw = rec + SLK_REC_HEADER;
    uint32_t cell = 0;
    for(int d = 0; d < c; ++d) {
        const uint32_t person = (cw[d >> 1] >> (16 * (d & 1))) & 0xffffu;
        cell |= (uint32_t) tm.pmk[person] << (2 * d);
    }
    const uint32_t s = cell & 3u;
    const uint32_t peel_lm = (uint32_t) tm.gc[peel] & 15u;
    double tc[4];
    slk_load_prior<false>(pg, tm, peel, tc);
    double t = slk_sel4(tc, g);
    if((rec[3] & 1u) && !((peel_lm >> g) & 1u)) t = 0.0;
    const uint32_t* pw = cw + ((c + 1) >> 1);
    const uint32_t* kw = pw + SLK_REC_PREV * nprev;
#define SLK_DIGC(pos) ((cell >> (2u * (pos))) & 3u)
    if(type == SLK_CHILD_PEEL) {
        const uint32_t k0 = kw[0];
        t *= slk_trans_prob(sex_linked, (k0 >> 28) & 1u, SLK_DIGC((k0 >> 20) & 15u), SLK_DIGC((k0 >> 24) & 15u), g,
                            tm.tables + 8 * (peel - E.F));
    }
    // a genotype the elimination ruled out has t == 0 and an unwritten (stale but finite) cell: 0 * x == 0
    for(int jp = 0; jp < nprev; ++jp) t *= *slk_backward_cell(tm, pw + SLK_REC_PREV * jp, cell, s, g);
    if(type == SLK_PARENT_PEEL) {
        double cp = 1.0;
        for(int k = 0; k < nkids; ++k) {
            const uint32_t kd = kw[k];
            const uint32_t mp = (kd >> 20) & 15u, fp = (kd >> 24) & 15u;
            const uint32_t m = (mp == SLK_POS_PEEL) ? g : SLK_DIGC(mp);
            const uint32_t f = (fp == SLK_POS_PEEL) ? g : SLK_DIGC(fp);
            const uint32_t kg = SLK_DIGC((kd >> 16) & 15u);
            cp *= slk_trans_prob(sex_linked, (kd >> 28) & 1u, m, f, kg, tm.tables + 8 * ((int)(kd & 0xffffu) - E.F));
        }
        t *= cp;
    }
#undef SLK_DIGC
    return t;
}

// SamplerRfunction::sample (sampler_rfunction.cc:159-188) given the four presum values and the draw:
// normalise (rfunction.cc:200-209), inverse CDF with strict `r < cumulative`, fall back to the last non-zero entry
SLK_DEV uint32_t slk_sample4(double dd[4], double r) {
    const double total = dd[0] + dd[1] + dd[2] + dd[3];
    if(total != 0.0) {
#pragma unroll
        for(int k = 0; k < 4; ++k) dd[k] /= total;
    }
    double cum = 0.0;
    int last = 0, chosen = -1;
#pragma unroll
    for(int k = 0; k < 4; ++k) {
        cum += dd[k];
        if(chosen < 0) {
            if(r < cum) chosen = k;
            else if(dd[k] != 0.0) last = k;
        }
    }
    return (uint32_t)(chosen >= 0 ? chosen : last);
}

// per-child transmission tables (sampler_rfunction.cc:102-157): for parent genotype UU / AA the transmitted allele
// is certain; for AU / UA it follows the neighbouring meioses.  q = 2 * (person - F) + parent.
SLK_DEV void slk_stage_transmission(const SlkEnv& E, const SlkTeam& tm, int q, bool has_left, bool has_right) {
    const int person = E.F + (q >> 1), parent = q & 1;
    double tmp0 = 0.5, tmp1 = 0.5;
    if(has_left) {
        const bool cross = ((tm.dgl[person] >> parent) & 1u) != 0;
        tmp0 *= cross ? tm.scal[2] : tm.scal[3];
        tmp1 *= cross ? tm.scal[3] : tm.scal[2];
    }
    if(has_right) {
        const bool cross = ((tm.dgr[person] >> parent) & 1u) != 0;
        tmp0 *= cross ? tm.scal[0] : tm.scal[1];
        tmp1 *= cross ? tm.scal[1] : tm.scal[0];
    }
    const double total = tmp0 + tmp1;
    const double u_au = tmp1 / total, u_ua = tmp0 / total;
    double* tb = tm.tables + 8 * (person - E.F) + 4 * parent;
    tb[0] = u_au; tb[1] = 1.0 - u_au;
    tb[2] = u_ua; tb[3] = 1.0 - u_ua;
}

// trait_prob x recombination probability per child and (i, j) (trait_rfunction.cc:9-22); q = 4 * (person - F) + 2 i + j
SLK_DEV void slk_stage_trait_weight(const SlkEnv& E, const SlkTeam& tm, int q, double trait_prob) {
    const bool sex_linked = E.sex_linked != 0;
    const int person = E.F + (q >> 2);
    const uint32_t i = (q >> 1) & 1u, j = q & 1u;
    const uint32_t l = tm.dgl[person], r = tm.dgr[person];
    double t = 1.0;
    t *= ((l & 1u) == i) ? tm.scal[1] : tm.scal[0];
    t *= ((r & 1u) == i) ? tm.scal[3] : tm.scal[2];
    if(!sex_linked) {
        t *= (((l >> 1) & 1u) == j) ? tm.scal[1] : tm.scal[0];
        t *= (((r >> 1) & 1u) == j) ? tm.scal[3] : tm.scal[2];
    }
    tm.tables[q] = trait_prob * t;
}

// meiosis indicators of one non-founder given the sampled genotypes (locus_sampler2.cc:32-125); returns the byte
// written to the descent graph.  `draw(parent)` supplies the uniform of a homozygous parent's Bernoulli.
template<class Draw>
SLK_DEV uint32_t slk_sample_indicators(const SlkEnv& E, const SlkTeam& tm, int i, int mother, int father,
                                       bool use_left, bool use_right, double th_left, double th_right, Draw draw) {
    const bool sex_linked = E.sex_linked != 0;
    const uint32_t trait = tm.pmk[i];
    uint32_t out = 0;
#pragma unroll
    for(int parent = 0; parent < 2; ++parent) {
        if(parent == 1 && sex_linked) break;                        // paternal := 0 (:115-118)
        const int pid = parent == 0 ? mother : father;
        const uint32_t pt = tm.pmk[pid];
        // allele the kid received from this parent: U = 0, A = 1
        const uint32_t allele = parent == 0 ? ((trait == SLK_UU || trait == SLK_UA) ? 0u : 1u)
                                            : ((trait == SLK_UU || trait == SLK_AU) ? 0u : 1u);
        uint32_t mi;
        if(pt >= 2u) {
            // heterozygous parent: forced (:32-39)
            mi = (allele == 0) ? ((pt == SLK_UA) ? 0u : 1u) : ((pt == SLK_UA) ? 1u : 0u);
        }
        else {
            double p0 = 1.0, p1 = 1.0;                              // :44-65
            if(use_left) {
                const uint32_t v = (tm.dgl[i] >> parent) & 1u;
                const double ith = 1.0 - th_left;
                p0 *= (v == 0) ? ith : th_left;
                p1 *= (v == 1) ? ith : th_left;
            }
            if(use_right) {
                const uint32_t v = (tm.dgr[i] >> parent) & 1u;
                const double ith = 1.0 - th_right;
                p0 *= (v == 0) ? ith : th_right;
                p1 *= (v == 1) ? ith : th_right;
            }
            mi = (draw(parent) < p0 / (p0 + p1)) ? 0u : 1u;
        }
        out |= mi << parent;
    }
    return out;
}

#endif
