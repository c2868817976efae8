// slk_kernels.cuh -- sm_100a kernels for the peeling hot path: the block-Gibbs locus sampler
// (L-sampler) and LOD scoring, both driven by the flattened peel program of slk_types.h.
//
// Execution model (B200: 148 SMs, 227 KB shared memory per CTA, FP64 on the CUDA cores):
//   * a TEAM of T threads (T = 32..512, chosen per plan) owns one unit of work at a time --
//     one marker locus for the sampler, one (interval, position) for LOD scoring -- and keeps
//     every peel matrix of that unit in its shared-memory arena; persistent CTAs stride over
//     the units, grid = SMs x resident CTAs;
//   * inside a unit the peel ops are walked level by level (dependency levels of the peel
//     forest); the valid cells of all ops of a level form one index space that the whole team
//     strides over, four cells (sixteen independent product chains) per thread, so a level of
//     many tiny ops costs one pass;
//   * only VALID cells are enumerated: the valid set of an op is a Cartesian product of
//     per-person legal genotype sets (4-bit masks), so the j-th valid cell is a mixed-radix
//     decode of j -- no per-locus index lists (the reference's matrix_indices[locus] tables,
//     peel_sequence_generator.cc:141-159, are what stops it scaling to 10k loci);
//   * a previous function is read by gathering 2-bit digits with precomputed shift/mask runs;
//   * the presum matrix is never stored: the backward pass recomputes the four candidates of
//     each op at the already sampled cutset assignment, four lanes per op, ops grouped into
//     backward levels; draws come from Philox keyed by (chain, iteration, locus, slot).
//
// Arithmetic follows the reference's operation order (compiled with -fmad=false) so peel
// matrices are bit-identical to the CPU R-functions: rfunction.cc:116-142,
// sampler_rfunction.cc:102-157,190-287,350-419, trait_rfunction.cc:9-137.
#ifndef SLK_KERNELS_CUH
#define SLK_KERNELS_CUH

#ifndef SLK_TILE_NS
#define SLK_TILE_NS 2          // slots (values of cutset digit 0) per thread in throughput mode: 2 keeps the
                               // kernels under 85 registers so that three 256-thread teams fit on an SM
#endif
#ifndef SLK_LS_MAXTHREADS
#define SLK_LS_MAXTHREADS 768
#endif

#include <stdint.h>
#include <float.h>

#include "swiftlink_b200.h"
#include "slk_types.h"
#include "slk_philox.cuh"

struct SlkLaunch {
    SlkDevPlan plan;
    uint8_t* dgp;                // [M][N] descent graph, bit0 maternal, bit1 paternal
    double* lod;                 // [(M-1)*nlod] log-sum accumulators
    double* gscratch;            // global arena slab: teams x (arena_doubles - smem_doubles)
    int* err;                    // [0] first error code, [1] unit
    uint64_t seed;
    uint64_t iteration;
    uint32_t chain;
    int window, offset;          // sampler units: locus = offset + unit * window
    int nunits;
    int ignore_left, ignore_right;
    int si_start;                // >= 0: sequential imputation walk from this locus (one team, loci in sequence)
    int si_batch;                // > 0: that many independent walks, one per team (graph g in rows [g M, (g + 1) M) of dgp)
    const int* si_starts;        // [si_batch] start locus of each walk
    int no_dg;                   // trait peel without a descent graph (P(T))
    int accumulate;              // LOD: 1 = log-sum into lod[], 0 = write dump_result/dump_prob
    int unit_base;               // LOD: first unit (debug: interval * nlod)
    int dump_k;                  // LOD debug: position whose matrices go to dump_mat
    double* dump_mat;            // dense matrices (pre-zeroed by the host)
    double* dump_pre;            // dense presum matrices
    double* dump_dist4;          // [nops][4]
    int* dump_pmk;               // [N]
    double* dump_result;         // sampler: [1]; LOD: [nunits]
    double* dump_prob;           // LOD: [nunits]
    long long* trace;            // optional: clock64() stamps of team 0's first unit (tuning aid)
    // batched replicates (ELOD, elod.cc:55-66): the graph is `nunits` short chromosomes of `period` loci laid
    // end to end; row r of the graph uses row r % period of the plan tables and has no neighbour across a
    // chromosome boundary.  LOD scoring then peels one interval per chromosome between its rows lod_row0 and
    // lod_row1 with the (two-locus) trait plan's interval 0.
    int period;
    int lod_row0, lod_row1;
    long long row_base;          // added to the graph row in the Philox key (replicates are sampled in chunks)
};

__constant__ uint8_t c_glist[16] = {
    // ascending list of the genotypes present in a 4-bit mask, 2 bits each
    0x00, 0x00, 0x01, 0x04, 0x02, 0x08, 0x09, 0x24, 0x03, 0x0C, 0x0D, 0x34, 0x0E, 0x38, 0x39, 0xE4
};

__constant__ double c_prior[5][4] = {
    {1.0, 0.0, 0.0, 0.0},
    {0.0, 1.0, 0.0, 0.0},
    {0.0, 0.0, 0.5, 0.5},
    {0.25, 0.25, 0.25, 0.25},
    {0.5, 0.5, 0.0, 0.0}
};

// ---- small device helpers ------------------------------------------------------------

__device__ __forceinline__ double lds_f64(uint32_t saddr) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(saddr));
    return v;
}

template<int T>
__device__ __forceinline__ void team_sync(int team) {
    if(T == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" :: "r"(team + 1), "r"(T) : "memory");
}

struct MatRef {
    uint32_t saddr;
    const double* gptr;
    bool in_smem;
};

__device__ __forceinline__ double mat_load(const MatRef& m, uint32_t idx) {
    if(m.in_smem) return lds_f64(m.saddr + idx * 8u);
    return m.gptr[idx];
}

struct TeamMem {
    double* arena;               // shared part
    uint32_t arena_saddr;
    double* garena;              // global part (offsets >= smem_doubles)
    int smem_doubles;
    double* tables;              // (N-F)*4
    double* scal;                // 16
    unsigned long long* lmw;     // nops
    double* ru;                  // nops uniform draws of the current locus (sampler)
    uint8_t* gc;                 // gcode row of this locus
    uint8_t* dgl;                // descent graph at the left neighbour / interval start
    uint8_t* dgr;                // right neighbour / interval end
    uint8_t* pmk;                // sampled genotypes
    int* red;
    const double* dprob;         // [N][4] disease probabilities (global)
};

__device__ __forceinline__ MatRef mat_ref(const TeamMem& tm, int off) {
    MatRef m;
    m.in_smem = off < tm.smem_doubles;
    m.saddr = tm.arena_saddr + (uint32_t) off * 8u;
    m.gptr = tm.garena + (off - tm.smem_doubles);
    return m;
}

__device__ __forceinline__ void mat_store(const TeamMem& tm, int off, double v) {
    if(off < tm.smem_doubles) tm.arena[off] = v;
    else tm.garena[off - tm.smem_doubles] = v;
}

__device__ __forceinline__ uint32_t digit(uint32_t cell, uint32_t pos) { return (cell >> (2 * pos)) & 3u; }

__device__ __forceinline__ uint32_t nvalid_of(unsigned long long lmw, int c) {
    uint32_t n = 1;
    for(int d = 0; d < c; ++d) n *= (uint32_t) __popc((uint32_t)(lmw >> (4 * d)) & 15u);
    return n;
}

// j-th valid cell: mixed-radix decode of j over the per-digit legal sets
__device__ __forceinline__ uint32_t decode_cell(uint32_t j, unsigned long long lmw, int c) {
    uint32_t cell = 0;
    for(int d = 0; d < c; ++d) {
        uint32_t lm = (uint32_t)(lmw >> (4 * d)) & 15u;
        uint32_t k = (uint32_t) __popc(lm);
        uint32_t q, r;
        if(k == 4)      { q = j >> 2; r = j & 3u; }
        else if(k == 1) { q = j; r = 0; }
        else if(k == 2) { q = j >> 1; r = j & 1u; }
        else            { q = __umulhi(j, 0x55555556u); r = j - 3u * q; }
        cell |= ((uint32_t)(c_glist[lm] >> (2 * r)) & 3u) << (2 * d);
        j = q;
    }
    return cell;
}

__device__ __forceinline__ unsigned long long full_lmw(int c) {
    return (c >= 16) ? ~0ull : ((1ull << (4 * c)) - 1ull);
}

__device__ __forceinline__ double sel4(const double v[4], uint32_t i) {
    double a = (i & 1u) ? v[1] : v[0];
    double b = (i & 1u) ? v[3] : v[2];
    return (i & 2u) ? b : a;
}

// sampler_rfunction.cc:350-419: one entry of the 64-entry transmission table, from the
// per-child tables staged at the start of the locus: tb[2*pg + a] = P(maternal allele a | mother
// genotype pg), tb[8 + 2*pg + a] the paternal one (sampler_rfunction.cc:102-157).  Under X
// linkage a heterozygous father transmits nothing (both entries 0) and a son copies the
// maternal allele (paternal entries of a homozygous father are 1, hetero kid genotypes are 0).
__device__ __forceinline__ double trans_prob(bool sex_linked, bool male, uint32_t m, uint32_t f, uint32_t kg,
                                             const double* tb) {
    const uint32_t am = ((kg >> 1) ^ kg) & 1u;       // maternal allele of the kid: AA, AU carry A
    const uint32_t ap = kg & 1u;                     // paternal allele: AA, UA carry A
    if(sex_linked && male && kg >= 2u) return 0.0;
    return tb[2u * m + am] * tb[8u + 2u * f + ap];
}

// rfunction.h:81-107
__device__ __forceinline__ bool affected_trait(uint32_t pt, uint32_t allele) {
    if(pt == SLK_UU) return false;
    if(pt == SLK_AA) return true;
    return (pt == SLK_AU) ? (allele == 0) : (allele == 1);
}

// rfunction.cc:94-113
__device__ __forceinline__ uint32_t phased_trait(bool sex_linked, bool male, uint32_t m, uint32_t f,
                                                 uint32_t i, uint32_t j) {
    bool ma = affected_trait(m, i), pa = affected_trait(f, j);
    if(sex_linked && male) return ma ? SLK_AA : SLK_UU;
    if(ma) return pa ? SLK_AA : SLK_AU;
    return pa ? SLK_UA : SLK_UU;
}

// trait_rfunction.cc:106-127: sum, in (i, j) order, of the recombination weights of the
// transmitted-allele choices that produce the kid's genotype.  The choices factorise: the kid's
// genotype fixes whether the maternal (paternal) allele must be the A allele, and the parent's
// genotype fixes which of its two alleles (i = 0, 1) qualify -- a 2-bit mask per parent, read
// from a packed table indexed by (parent genotype, required) [rfunction.h:81-107].
__device__ __forceinline__ double trait_child_sum(bool sex_linked, bool male, uint32_t m, uint32_t f,
                                                  uint32_t kg, const double* w) {
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

struct Prog {
    const uint32_t* stream;
    const uint16_t* op_start;
    const uint16_t* lops;
    const uint32_t* lpf;
    const uint32_t* flevel_quads;
    const uint16_t* flevel_start;
    const uint16_t* bops;
    const uint16_t* blevel_start;
};

// copies the program into shared memory (whole CTA), returns the shared views
__device__ __forceinline__ Prog stage_program(const SlkProgram& p, int nops, unsigned char* smem) {
    uint32_t* s_stream = (uint32_t*) smem;
    uint32_t* s_lpf = s_stream + p.stream_words;
    uint32_t* s_fq = s_lpf + nops;
    uint16_t* s_op_start = (uint16_t*)(s_fq + p.n_flevels);
    uint16_t* s_lops = s_op_start + nops;
    uint16_t* s_fl = s_lops + nops;
    uint16_t* s_bops = s_fl + (p.n_flevels + 1);
    uint16_t* s_bl = s_bops + nops;
    for(int i = threadIdx.x; i < p.stream_words; i += blockDim.x) s_stream[i] = p.stream[i];
    for(int i = threadIdx.x; i < nops; i += blockDim.x) {
        s_lpf[i] = p.lpf[i]; s_op_start[i] = p.op_start[i]; s_lops[i] = p.lops[i]; s_bops[i] = p.bops[i];
    }
    for(int i = threadIdx.x; i < p.n_flevels; i += blockDim.x) s_fq[i] = p.flevel_quads[i];
    for(int i = threadIdx.x; i <= p.n_flevels; i += blockDim.x) s_fl[i] = p.flevel_start[i];
    for(int i = threadIdx.x; i <= p.n_blevels; i += blockDim.x) s_bl[i] = p.blevel_start[i];
    __syncthreads();
    Prog g;
    g.stream = s_stream; g.lpf = s_lpf; g.flevel_quads = s_fq; g.op_start = s_op_start; g.lops = s_lops;
    g.flevel_start = s_fl; g.bops = s_bops; g.blevel_start = s_bl;
    return g;
}

__device__ __forceinline__ TeamMem carve_team(const SlkLaunch& L, const SlkProgram& p, unsigned char* base, int gteam) {
    const SlkDevPlan& P = L.plan;
    SlkTeamLayout lay = slk_team_layout(P.N, P.F, P.nops, p.smem_doubles, p.table_doubles_per_child);
    TeamMem tm;
    tm.arena = (double*)(base + lay.arena);
    tm.arena_saddr = (uint32_t) __cvta_generic_to_shared(tm.arena);
    tm.smem_doubles = p.smem_doubles;
    tm.garena = L.gscratch + (size_t) gteam * (size_t)(p.arena_doubles - p.smem_doubles);
    tm.tables = (double*)(base + lay.tables);
    tm.scal = (double*)(base + lay.scal);
    tm.lmw = (unsigned long long*)(base + lay.lmw);
    tm.ru = (double*)(base + lay.ru);
    tm.gc = base + lay.bytes;
    tm.dgl = tm.gc + lay.nbytes;
    tm.dgr = tm.dgl + lay.nbytes;
    tm.pmk = tm.dgr + lay.nbytes;
    tm.red = (int*)(base + lay.red);
    tm.dprob = P.disease_prob;
    return tm;
}

// marker prior of person `peel` at the staged locus (person.cc:224-245 via prior classes)
__device__ __forceinline__ void load_prior(const TeamMem& tm, int person, double tc[4]) {
    uint32_t cls = tm.gc[person] >> 4;
    if(cls == SLK_PRIOR_PERSON) {
        // ELOD's simulated trait locus: the person's disease probabilities (person.h:204-208)
#pragma unroll
        for(int g = 0; g < 4; ++g) tc[g] = tm.dprob[4 * person + g];
        return;
    }
#pragma unroll
    for(int g = 0; g < 4; ++g) tc[g] = tm.scal[16 + cls * 4 + g];      // classes 0-4 fixed, 5-6 founder priors of this locus
}

// ---- one forward tile, executed by one thread -----------------------------------------------
//
// A tile is one valid assignment of cutset digits 1..c-1 (a "row") times the four values of
// cutset digit 0 (slots) times the four genotypes of the peel node (variants): 16 product
// chains per thread.  The gather index of a previous function is computed once per tile; a
// previous function that is not keyed on digit 0 (or not on the peel node) is loaded once per
// slot group instead of 16 times.
//
// TRAIT = false: SamplerRfunction::evaluate_element, TRAIT = true: TraitRfunction::evaluate_element
// DBG adds the dense parity dumps (separate instantiation, keeps the hot kernels lean).

// applies one previous function to the 16 products; SM selects shared vs global arena loads
template<int NS, bool SM, bool PERM>
__device__ __forceinline__ void apply_prev(double (&tmp)[NS][4], const bool (&act)[NS],
                                           const uint32_t (&gvp)[NS], uint32_t saddr, const double* gptr,
                                           uint32_t base, uint32_t d0shift, uint32_t pshift, uint32_t s0) {
    auto ld = [&](uint32_t i) -> double { return SM ? lds_f64(saddr + i * 8u) : gptr[i]; };
    const bool has0 = d0shift != SLK_NO_SHIFT, hasp = pshift != SLK_NO_SHIFT;
    if(!hasp) {
        if(!has0) {
            // same cell for the whole tile
            bool any = false;
#pragma unroll
            for(int s = 0; s < NS; ++s) any = any || act[s];
            if(any) {
                const double x = ld(base);
#pragma unroll
                for(int s = 0; s < NS; ++s)
#pragma unroll
                    for(int v = 0; v < 4; ++v) tmp[s][v] *= x;
            }
        }
        else {
#pragma unroll
            for(int s = 0; s < NS; ++s) {
                if(act[s]) {
                    const double x = ld(base + ((s0 + (uint32_t) s) << d0shift));
#pragma unroll
                    for(int v = 0; v < 4; ++v) tmp[s][v] *= x;
                }
            }
        }
    }
    else if(!has0 && !PERM) {
        // keyed on the peel node only: one load per genotype, shared by the four slots
#pragma unroll
        for(int v = 0; v < 4; ++v) {
            bool any = false;
#pragma unroll
            for(int s = 0; s < NS; ++s) any = any || (tmp[s][v] != 0.0);
            if(any) {
                const double x = ld(base + ((uint32_t) v << pshift));
#pragma unroll
                for(int s = 0; s < NS; ++s) tmp[s][v] *= x;
            }
        }
    }
    else {
#pragma unroll
        for(int s = 0; s < NS; ++s) {
            const uint32_t bs = base + (has0 ? ((s0 + (uint32_t) s) << d0shift) : 0u);
#pragma unroll
            for(int v = 0; v < 4; ++v) {
                if(tmp[s][v] != 0.0) {
                    const uint32_t g = PERM ? ((gvp[s] >> (2 * v)) & 3u) : (uint32_t) v;
                    tmp[s][v] *= ld(bs + (g << pshift));
                }
            }
        }
    }
}

// NS = 4: the thread owns the whole tile (throughput mode, big levels); NS = 1: the thread owns
// slot s0 only (latency mode: a level with few rows is spread over four times as many threads)
template<bool TRAIT, bool DBG, int NS>
__device__ __forceinline__ void run_tile(const SlkLaunch& L, const Prog& pg, const TeamMem& tm, uint32_t op, uint32_t row, uint32_t s0) {
    const SlkDevPlan& P = L.plan;
    const bool sex_linked = P.sex_linked != 0;
    const uint32_t* rec = pg.stream + pg.op_start[op];
    const uint32_t w0 = rec[0];
    const int type = w0 & 7u, c = (w0 >> 4) & 15u, nprev = (w0 >> 8) & 15u, nkids = (w0 >> 12) & 15u;
    const int peel = w0 >> 16;
    const int mat_off = (int) rec[1];
    const unsigned long long lmw = TRAIT ? ((unsigned long long) rec[4] | ((unsigned long long) rec[5] << 32)) : tm.lmw[op];
    const int ch = c > 0 ? c - 1 : 0;
    const unsigned long long lmw_hi = lmw >> 4;
    const uint32_t nrows = nvalid_of(lmw_hi, ch);
    if(row >= nrows) return;
    const uint32_t lm0 = c > 0 ? ((uint32_t) lmw & 15u) : 1u;             // legal values of digit 0 (c = 0: one cell)
    const uint32_t rowcell = ((lmw_hi == full_lmw(ch)) ? row : decode_cell(row, lmw_hi, ch)) << 2;
    const uint32_t peel_lm = TRAIT ? (rec[6] & 15u) : ((uint32_t) tm.gc[peel] & 15u);
    const bool peel_in_prev = (rec[3] & 1u) != 0;
    const uint32_t* pw = rec + SLK_REC_HEADER + ((c + 1) >> 1);
    const uint32_t* kw = pw;
    for(int jp = 0; jp < nprev; ++jp) kw += 2 + ((((kw[1] >> 8) & 0xffu) + 1) >> 1);
    const bool perm = TRAIT && type == SLK_CHILD_PEEL;

    double tc[4];
    if(TRAIT) {
#pragma unroll
        for(int g = 0; g < 4; ++g) tc[g] = P.disease_prob[peel * 4 + g];
    }
    else load_prior(tm, peel, tc);

    bool act[NS];
    uint32_t gvp[NS];
    double tmp[NS][4];
    // digit value of cutset position `pos` in slot s
#define SLK_DIG(pos, s) (((pos) == 0u) ? (s0 + (uint32_t)(s)) : ((rowcell >> (2u * (pos))) & 3u))

    // ---- initial factor: prior of the peel-node genotype (x transmission for a sampler child peel)
    const uint32_t k0 = (type == SLK_CHILD_PEEL) ? kw[0] : 0u;
    const uint32_t mpos0 = (k0 >> 20) & 15u, fpos0 = (k0 >> 24) & 15u;
    const bool male0 = (k0 >> 28) & 1u;
#pragma unroll
    for(int s = 0; s < NS; ++s) {
        act[s] = (lm0 >> (s0 + (uint32_t) s)) & 1u;
        gvp[s] = 0xE4u;
        if(perm) {
            // variants are the four (maternal, paternal) transmitted-allele choices (trait_rfunction.cc:44-66)
            const uint32_t m = SLK_DIG(mpos0, s), f = SLK_DIG(fpos0, s);
            uint32_t g = 0;
#pragma unroll
            for(uint32_t ij = 0; ij < 4; ++ij) g |= phased_trait(sex_linked, male0, m, f, ij >> 1, ij & 1u) << (2 * ij);
            gvp[s] = g;
        }
#pragma unroll
        for(int v = 0; v < 4; ++v) {
            const uint32_t g = perm ? ((gvp[s] >> (2 * v)) & 3u) : (uint32_t) v;
            double t = act[s] ? (perm ? sel4(tc, g) : tc[v]) : 0.0;
            if(!TRAIT) {
                // a previous function keyed on the peel node is 0 wherever that genotype was
                // eliminated (its cell is not in valid_indices): the product is 0
                if(peel_in_prev && !((peel_lm >> v) & 1u)) t = 0.0;
                if(type == SLK_CHILD_PEEL && t != 0.0) {
                    const uint32_t m = SLK_DIG(mpos0, s), f = SLK_DIG(fpos0, s);
                    t *= trans_prob(sex_linked, male0, m, f, (uint32_t) v, tm.tables + 16 * (peel - P.F));
                }
            }
            tmp[s][v] = t;
        }
    }

    // ---- previous functions, in the reference's order
    for(int jp = 0; jp < nprev; ++jp) {
        const int poff = (int) pw[0];
        const uint32_t meta = pw[1];
        const uint32_t pshift = meta & 0xffu, nruns = (meta >> 8) & 0xffu, d0shift = (meta >> 16) & 0xffu;
        uint32_t base = 0;
        for(uint32_t r = 0; r < nruns; ++r) {
            const uint32_t run = (pw[2 + (r >> 1)] >> (16 * (r & 1u))) & 0xffffu;
            base |= ((rowcell >> (run & 31u)) & ((1u << (run >> 10)) - 1u)) << ((run >> 5) & 31u);
        }
        const bool in_smem = poff < tm.smem_doubles;
        const uint32_t saddr = tm.arena_saddr + (uint32_t) poff * 8u;
        const double* gptr = tm.garena + (poff - tm.smem_doubles);
        if(perm) {
            if(in_smem) apply_prev<NS, true, true>(tmp, act, gvp, saddr, gptr, base, d0shift, pshift, s0);
            else        apply_prev<NS, false, true>(tmp, act, gvp, saddr, gptr, base, d0shift, pshift, s0);
        }
        else {
            if(in_smem) apply_prev<NS, true, false>(tmp, act, gvp, saddr, gptr, base, d0shift, pshift, s0);
            else        apply_prev<NS, false, false>(tmp, act, gvp, saddr, gptr, base, d0shift, pshift, s0);
        }
        pw += 2 + ((nruns + 1) >> 1);
    }

    // ---- transmission to the children / recombination weights
    if(TRAIT && type == SLK_CHILD_PEEL) {
        const double* w = tm.tables + 4 * (peel - P.F);
#pragma unroll
        for(int s = 0; s < NS; ++s)
#pragma unroll
            for(int v = 0; v < 4; ++v) tmp[s][v] *= w[v];
    }
    else if(type == SLK_PARENT_PEEL) {
        // child_prob = prod_k T_k (sampler_rfunction.cc:258-279 / trait_rfunction.cc:91-129).  A child
        // whose own, mother's and father's digits are all outside cutset digit 0 has the same factor
        // in the four slots of the tile: it is evaluated once per peel genotype, not sixteen times.
#pragma unroll
        for(int v = 0; v < 4; ++v) {
            bool any = false;
#pragma unroll
            for(int s = 0; s < NS; ++s) any = any || (tmp[s][v] != 0.0);
            if(!any) continue;
            double cp[NS];
#pragma unroll
            for(int s = 0; s < NS; ++s) cp[s] = 1.0;
            for(int k = 0; k < nkids; ++k) {
                const uint32_t kd = kw[k];
                const uint32_t mp = (kd >> 20) & 15u, fp = (kd >> 24) & 15u, kp = (kd >> 16) & 15u;
                const bool male = (kd >> 28) & 1u;
                const double* tb = tm.tables + (TRAIT ? 4 : 16) * ((int)(kd & 0xffffu) - P.F);
                if(NS > 1 && mp != 0u && fp != 0u && kp != 0u) {
                    const uint32_t m = (mp == SLK_POS_PEEL) ? (uint32_t) v : ((rowcell >> (2u * mp)) & 3u);
                    const uint32_t f = (fp == SLK_POS_PEEL) ? (uint32_t) v : ((rowcell >> (2u * fp)) & 3u);
                    const uint32_t kg = (rowcell >> (2u * kp)) & 3u;
                    const double t = TRAIT ? trait_child_sum(sex_linked, male, m, f, kg, tb)
                                           : trans_prob(sex_linked, male, m, f, kg, tb);
#pragma unroll
                    for(int s = 0; s < NS; ++s) cp[s] *= t;
                }
                else {
#pragma unroll
                    for(int s = 0; s < NS; ++s) {
                        const uint32_t m = (mp == SLK_POS_PEEL) ? (uint32_t) v : SLK_DIG(mp, s);
                        const uint32_t f = (fp == SLK_POS_PEEL) ? (uint32_t) v : SLK_DIG(fp, s);
                        const uint32_t kg = SLK_DIG(kp, s);
                        cp[s] *= TRAIT ? trait_child_sum(sex_linked, male, m, f, kg, tb)
                                       : trans_prob(sex_linked, male, m, f, kg, tb);
                    }
                }
            }
#pragma unroll
            for(int s = 0; s < NS; ++s) tmp[s][v] *= cp[s];
        }
    }

    // ---- the sum over the peel node
    const int dense_off = (int) rec[2];
#pragma unroll
    for(int s = 0; s < NS; ++s) {
        if(!act[s]) continue;
        double total = 0.0;
#pragma unroll
        for(int v = 0; v < 4; ++v) total += tmp[s][v];
        const uint32_t cell = rowcell | (s0 + (uint32_t) s);
        mat_store(tm, mat_off + (int) cell, total);
        if(DBG) {
            if(L.dump_mat) L.dump_mat[dense_off + cell] = total;
            if(!TRAIT && L.dump_pre) {
#pragma unroll
                for(int v = 0; v < 4; ++v) L.dump_pre[4 * (size_t) dense_off + cell + ((size_t) v << (2 * c))] = tmp[s][v];
            }
        }
    }
#undef SLK_DIG
}

// ---- one term of the backward pass: presum(cell, g) of op recomputed (sampler only) --------
__device__ __forceinline__ double sampler_term(const SlkLaunch& L, const Prog& pg, const TeamMem& tm,
                                               const uint32_t* rec, uint32_t cell, uint32_t g) {
    const SlkDevPlan& P = L.plan;
    const bool sex_linked = P.sex_linked != 0;
    const uint32_t w0 = rec[0];
    const int type = w0 & 7u, c = (w0 >> 4) & 15u, nprev = (w0 >> 8) & 15u, nkids = (w0 >> 12) & 15u;
    const int peel = w0 >> 16;
    const uint32_t peel_lm = (uint32_t) tm.gc[peel] & 15u;
    double tc[4];
    load_prior(tm, peel, tc);
    double t = sel4(tc, g);
    if((rec[3] & 1u) && !((peel_lm >> g) & 1u)) t = 0.0;
    const uint32_t* pw = rec + SLK_REC_HEADER + ((c + 1) >> 1);
    const uint32_t* kw = pw;
    for(int jp = 0; jp < nprev; ++jp) kw += 2 + ((((kw[1] >> 8) & 0xffu) + 1) >> 1);
    if(type == SLK_CHILD_PEEL && t != 0.0) {
        uint32_t k0 = kw[0];
        uint32_t m = digit(cell, (k0 >> 20) & 15u), f = digit(cell, (k0 >> 24) & 15u);
        t *= trans_prob(sex_linked, (k0 >> 28) & 1u, m, f, g, tm.tables + 16 * (peel - P.F));
    }
    for(int jp = 0; jp < nprev; ++jp) {
        const MatRef mr = mat_ref(tm, (int) pw[0]);
        const uint32_t meta = pw[1];
        const uint32_t pshift = meta & 0xffu, nruns = (meta >> 8) & 0xffu;
        const uint32_t d0shift = (meta >> 16) & 0xffu;
        uint32_t idx = 0;
        for(uint32_t r = 0; r < nruns; ++r) {
            uint32_t run = (pw[2 + (r >> 1)] >> (16 * (r & 1u))) & 0xffffu;
            uint32_t src = run & 31u, dst = (run >> 5) & 31u, mask = (1u << (run >> 10)) - 1u;
            idx |= ((cell >> src) & mask) << dst;
        }
        if(d0shift != SLK_NO_SHIFT) idx += (cell & 3u) << d0shift;
        if(pshift != SLK_NO_SHIFT) idx += g << pshift;
        if(pshift == SLK_NO_SHIFT || t != 0.0) t *= mat_load(mr, idx);
        pw += 2 + ((nruns + 1) >> 1);
    }
    if(type == SLK_PARENT_PEEL && t != 0.0) {
        double cp = 1.0;
        for(int k = 0; k < nkids; ++k) {
            uint32_t kd = kw[k];
            uint32_t mp = (kd >> 20) & 15u, fp = (kd >> 24) & 15u;
            uint32_t m = (mp == SLK_POS_PEEL) ? g : digit(cell, mp);
            uint32_t f = (fp == SLK_POS_PEEL) ? g : digit(cell, fp);
            uint32_t kg = digit(cell, (kd >> 16) & 15u);
            cp *= trans_prob(sex_linked, (kd >> 28) & 1u, m, f, kg, tm.tables + 16 * ((int)(kd & 0xffffu) - P.F));
        }
        t *= cp;
    }
    return t;
}

__device__ __forceinline__ void raise_error(int* err, int code, int unit) {
    if(atomicCAS(&err[0], 0, code) == 0) err[1] = unit;
}

// ============================================================================================
// L-sampler: LocusSampler::set_locus_minimal + step (locus_sampler2.cc:128-181) for every locus
// of one window class.  Replaces lsampler_kernel (cuda_lsampler.cu:403-449).
// ============================================================================================
template<int T, bool DBG>
__global__ void __launch_bounds__(SLK_LS_MAXTHREADS, 1) slk_lsampler_kernel(const SlkLaunch L) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SlkDevPlan& P = L.plan;
    const SlkProgram& pr = P.ls;
    const Prog pg = stage_program(pr, P.nops, smem);

    const int teams_per_cta = blockDim.x / T;
    const int team = threadIdx.x / T;
    const int tid = threadIdx.x % T;
    const int lane = tid & 31;
    const int gteam = blockIdx.x * teams_per_cta + team;
    const int total_teams = gridDim.x * teams_per_cta;
    const TeamMem tm = carve_team(L, pr, smem + pr.prog_smem_bytes + (size_t) team * pr.team_smem_bytes, gteam);
    const int N = P.N, F = P.F, M = P.M;
    const bool sex_linked = P.sex_linked != 0;

    // Batched sequential imputation (L.si_batch > 0): team g walks graph g -- M loci in sequence, rows [g M, (g + 1) M)
    // of the graph buffer, run number L.iteration + g, start locus L.si_starts[g] -- so that hundreds of the
    // reference's independent start_from runs (sequential_imputation.cc:47-115) share one launch.
    const int si_batch = L.si_batch;
    for(int graph = si_batch ? gteam : 0; graph < (si_batch ? si_batch : 1); graph += (si_batch ? total_teams : 1)) {
    uint8_t* const dgp = L.dgp + (si_batch ? (size_t) graph * (size_t) M * (size_t) N : (size_t) 0);
    const uint64_t iteration = L.iteration + (uint64_t)(si_batch ? graph : 0);
    const int si_start = si_batch ? L.si_starts[graph] : L.si_start;
    double* const si_weight = L.dump_result ? L.dump_result + (si_batch ? graph : 0) : (double*) 0;
    for(int unit = si_batch ? 0 : gteam; unit < L.nunits; unit += (si_batch ? 1 : total_teams)) {
        int locus = L.offset + unit * L.window;
        bool ign_left = L.ignore_left != 0, ign_right = L.ignore_right != 0;
        if(si_start >= 0) {
            // LocusSampler::start_from (locus_sampler2.cc:209-241): the start locus alone, then
            // leftwards conditioning on the right neighbour only, then rightwards on the left one
            if(unit == 0)               { locus = si_start; ign_left = true; ign_right = true; }
            else if(unit <= si_start) { locus = si_start - unit; ign_left = true; ign_right = false; }
            else                        { locus = unit; ign_left = false; ign_right = true; }
        }
        const int plocus = L.period ? locus % L.period : locus;          // row of the plan tables
        const bool has_left = (plocus != 0), has_right = (plocus != (L.period ? L.period : M) - 1);

        const bool tracing = L.trace != 0 && gteam == 0 && unit == gteam && tid == 0;
        int tstamp = 0;
        if(tracing) L.trace[tstamp++] = clock64();
        // ---- stage: genotype codes, neighbouring descent-graph rows, thetas, founder priors
        for(int i = tid; i < N; i += T) {
            tm.gc[i] = P.gcode[(size_t) plocus * N + i];
            tm.dgl[i] = has_left ? dgp[(size_t)(locus - 1) * N + i] : 0;
            tm.dgr[i] = has_right ? dgp[(size_t)(locus + 1) * N + i] : 0;
        }
        for(int i = tid; i < 28; i += T)
            tm.scal[16 + i] = (i < 20) ? c_prior[i >> 2][i & 3] : P.fprior[(size_t) plocus * 8 + (i - 20)];
        // the genotype draws of this locus, one per peel op (Philox keyed by chain, iteration, locus, op)
        for(int op = tid; op < P.nops; op += T)
            tm.ru[op] = slk_uniform(L.seed, L.chain, iteration, (uint32_t)(locus + L.row_base), (uint32_t) op);
        if(tid == 0) {
            // sampler_rfunction.h:84-100: theta2 (left interval) / theta (right interval)
            double th2 = 1.0, ath2 = 1.0, th = 1.0, ath = 1.0;
            if(has_left && !ign_left)   { th2 = P.theta[plocus - 1]; ath2 = 1.0 - th2; }
            if(has_right && !ign_right) { th = P.theta[plocus]; ath = 1.0 - th; }
            tm.scal[0] = th; tm.scal[1] = ath; tm.scal[2] = th2; tm.scal[3] = ath2;
        }
        team_sync<T>(team);

        // per-child transmission tables (sampler_rfunction.cc:102-157): for parent genotype UU / AA
        // the transmitted allele is certain; for AU / UA it follows the neighbouring meioses
        for(int q = tid; q < 2 * (N - F); q += T) {
            const int person = F + (q >> 1), parent = q & 1;
            double tmp0 = 0.5, tmp1 = 0.5;
            if(has_left) {
                bool cross = ((tm.dgl[person] >> parent) & 1u) != 0;
                tmp0 *= cross ? tm.scal[2] : tm.scal[3];
                tmp1 *= cross ? tm.scal[3] : tm.scal[2];
            }
            if(has_right) {
                bool cross = ((tm.dgr[person] >> parent) & 1u) != 0;
                tmp0 *= cross ? tm.scal[0] : tm.scal[1];
                tmp1 *= cross ? tm.scal[1] : tm.scal[0];
            }
            const double total = tmp0 + tmp1;
            const double u_au = tmp1 / total, u_ua = tmp0 / total;
            double* tb = tm.tables + 16 * (person - F) + 8 * parent;
            const bool xpat = sex_linked && parent == 1;
            const bool son = sex_linked && P.male[person];
            // homozygous parent
            tb[2 * SLK_UU + 0] = 1.0; tb[2 * SLK_UU + 1] = (xpat && son) ? 1.0 : 0.0;
            tb[2 * SLK_AA + 0] = (xpat && son) ? 1.0 : 0.0; tb[2 * SLK_AA + 1] = 1.0;
            // heterozygous parent
            tb[2 * SLK_AU + 0] = xpat ? 0.0 : u_au; tb[2 * SLK_AU + 1] = xpat ? 0.0 : 1.0 - u_au;
            tb[2 * SLK_UA + 0] = xpat ? 0.0 : u_ua; tb[2 * SLK_UA + 1] = xpat ? 0.0 : 1.0 - u_ua;
        }
        // per-op legal masks of the cutset at this locus
        for(int op = tid; op < P.nops; op += T) {
            const uint32_t* rec = pg.stream + pg.op_start[op];
            const int c = (rec[0] >> 4) & 15u;
            unsigned long long w = 0;
            for(int d = 0; d < c; ++d) {
                uint32_t person = (rec[SLK_REC_HEADER + (d >> 1)] >> (16 * (d & 1))) & 0xffffu;
                w |= (unsigned long long)(tm.gc[person] & 15u) << (4 * d);
            }
            tm.lmw[op] = w;
        }
        team_sync<T>(team);

        if(tracing) L.trace[tstamp++] = clock64();
        // ---- forward peel, level by level
        for(int lv = 0; lv < pr.n_flevels; ++lv) {
            const int b = pg.flevel_start[lv], e = pg.flevel_start[lv + 1];
            const uint32_t nq = pg.flevel_quads[lv];            // rows of this level
            if(nq * 4u <= 2u * T) {
                // latency mode: one (row, slot) per thread
                for(uint32_t q = tid; q < nq * 4u; q += T) {
                    const uint32_t row = q >> 2;
                    int lo = b, hi = e - 1;
                    while(lo < hi) { const int mid = (lo + hi + 1) >> 1; if(pg.lpf[mid] <= row) lo = mid; else hi = mid - 1; }
                    run_tile<false, DBG, 1>(L, pg, tm, pg.lops[lo], row - pg.lpf[lo], q & 3u);
                }
            }
            else {
                for(uint32_t q = tid; q < nq * (4u / SLK_TILE_NS); q += T) {
                    const uint32_t row = q / (4u / SLK_TILE_NS);
                    int lo = b, hi = e - 1;
                    while(lo < hi) { const int mid = (lo + hi + 1) >> 1; if(pg.lpf[mid] <= row) lo = mid; else hi = mid - 1; }
                    run_tile<false, DBG, SLK_TILE_NS>(L, pg, tm, pg.lops[lo], row - pg.lpf[lo], (q % (4u / SLK_TILE_NS)) * SLK_TILE_NS);
                }
            }
            team_sync<T>(team);
            if(tracing) L.trace[tstamp++] = clock64();
        }

        const uint32_t* last_rec = pg.stream + pg.op_start[P.last_op];
        const double result = mat_load(mat_ref(tm, (int) last_rec[1]), 0);
        if(si_weight && tid == 0) {
            if(si_start >= 0) si_weight[0] = (unit == 0 ? 0.0 : si_weight[0]) + log(result);   // SI weight
            else si_weight[0] = result;
        }
        if(result == 0.0) {
            // locus_sampler2.cc:137-142 exits the program; here the locus is left untouched
            if(tid == 0) raise_error(L.err, SLK_ERR_ZERO_LIKELIHOOD, locus);
            team_sync<T>(team);
            continue;
        }

        // ---- backward pass: SamplerRfunction::sample (sampler_rfunction.cc:159-188), 4 lanes per op
        for(int lv = 0; lv < pr.n_blevels; ++lv) {
            const int b = pg.blevel_start[lv], e = pg.blevel_start[lv + 1];
            // one op per warp (lanes 0-3 = the four candidate genotypes): ops of different shapes in
            // one warp would serialise each other's control flow on this latency-critical chain
            // When the level has no more ops than the team has warps every op gets its own warp (lanes
            // 0-3 = the four candidate genotypes): ops of different shapes packed in one warp would
            // serialise each other's control flow on this latency-critical chain.
            const bool spread = (e - b) <= T / 32;
            for(int base = b; base < e; base += (spread ? T / 32 : T / 4)) {
                const int q = base + (spread ? (tid >> 5) : (tid >> 2));
                const uint32_t g = tid & 3u;
                double d = 0.0;
                int op = 0, peel = 0;
                const bool live = q < e && (!spread || lane < 4);
                if(live) {
                    op = pg.bops[q];
                    const uint32_t* rec = pg.stream + pg.op_start[op];
                    const int c = (rec[0] >> 4) & 15u;
                    peel = rec[0] >> 16;
                    uint32_t cell = 0;
                    for(int k = 0; k < c; ++k) {
                        uint32_t person = (rec[SLK_REC_HEADER + (k >> 1)] >> (16 * (k & 1))) & 0xffffu;
                        cell |= (uint32_t) tm.pmk[person] << (2 * k);
                    }
                    d = sampler_term(L, pg, tm, rec, cell, g);
                }
                double dd[4];
                const int lbase = lane & ~3;
#pragma unroll
                for(int k = 0; k < 4; ++k) dd[k] = __shfl_sync(0xffffffffu, d, lbase + k);
                if(live && g == 0) {
                    if(DBG && L.dump_dist4) {
#pragma unroll
                        for(int k = 0; k < 4; ++k) L.dump_dist4[4 * op + k] = dd[k];
                    }
                    double total = dd[0] + dd[1] + dd[2] + dd[3];              // rfunction.cc:200-209
                    if(total != 0.0) {
#pragma unroll
                        for(int k = 0; k < 4; ++k) dd[k] /= total;
                    }
                    const double r = tm.ru[op];
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
                    tm.pmk[peel] = (uint8_t)(chosen >= 0 ? chosen : last);
                }
            }
            team_sync<T>(team);
            if(tracing) L.trace[tstamp++] = clock64();
        }

        // ---- meiosis indicators (locus_sampler2.cc:32-125)
        for(int i = F + tid; i < N; i += T) {
            const uint32_t trait = tm.pmk[i];
            uint32_t out = 0;
#pragma unroll
            for(int parent = 0; parent < 2; ++parent) {
                if(parent == 1 && sex_linked) break;                        // paternal := 0 (:115-118)
                const int pid = parent == 0 ? P.mother[i] : P.father[i];
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
                    if(has_left && !ign_left) {
                        const uint32_t v = (tm.dgl[i] >> parent) & 1u;
                        const double th = P.theta[plocus - 1], ith = 1.0 - th;
                        p0 *= (v == 0) ? ith : th;
                        p1 *= (v == 1) ? ith : th;
                    }
                    if(has_right && !ign_right) {
                        const uint32_t v = (tm.dgr[i] >> parent) & 1u;
                        const double th = P.theta[plocus], ith = 1.0 - th;
                        p0 *= (v == 0) ? ith : th;
                        p1 *= (v == 1) ? ith : th;
                    }
                    const double r = slk_uniform(L.seed, L.chain, iteration, (uint32_t)(locus + L.row_base),
                                                 (uint32_t)(P.nops + 2 * i + parent));
                    mi = (r < p0 / (p0 + p1)) ? 0u : 1u;
                }
                out |= mi << parent;
            }
            dgp[(size_t) locus * N + i] = (uint8_t) out;
        }
        if(tracing) L.trace[tstamp++] = clock64();
        if(DBG && L.dump_pmk) for(int i = tid; i < N; i += T) L.dump_pmk[i] = tm.pmk[i];
        if(si_start >= 0) __threadfence();          // the next locus of the walk reads the row just written
        team_sync<T>(team);
    }
    }
}

// ============================================================================================
// LOD scoring: Peeler::process (peeler.cc:79-103) with one team per (interval, position).
// Replaces lodscore_kernel (cuda_lodscore.cu:389-467).
// ============================================================================================
template<int T, bool DBG>
__global__ void __launch_bounds__(SLK_LS_MAXTHREADS, 1) slk_lodscore_kernel(const SlkLaunch L) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SlkDevPlan& P = L.plan;
    const SlkProgram& pr = P.lod;
    const Prog pg = stage_program(pr, P.nops, smem);

    const int teams_per_cta = blockDim.x / T;
    const int team = threadIdx.x / T;
    const int tid = threadIdx.x % T;
    const int lane = tid & 31;
    const int gteam = blockIdx.x * teams_per_cta + team;
    const int total_teams = gridDim.x * teams_per_cta;
    const TeamMem tm = carve_team(L, pr, smem + pr.prog_smem_bytes + (size_t) team * pr.team_smem_bytes, gteam);
    const int N = P.N, F = P.F;
    const bool sex_linked = P.sex_linked != 0;
    const double trait_prob = sex_linked ? 0.5 : 0.25;                      // trait_rfunction.cc:37,78

    for(int u = gteam; u < L.nunits; u += total_teams) {
        const int unit = L.unit_base + u;
        const int interval = L.period ? 0 : unit / P.nlod;
        const int k = L.period ? 0 : unit - interval * P.nlod;              // position k+1 of nlod
        const size_t row_l = L.period ? (size_t) unit * L.period + L.lod_row0 : (size_t) interval;
        const size_t row_r = L.period ? (size_t) unit * L.period + L.lod_row1 : (size_t) interval + 1;

        int ncross = 0;
        if(!L.no_dg) {
            for(int i = tid; i < N; i += T) {
                tm.dgl[i] = L.dgp[row_l * N + i];
                tm.dgr[i] = L.dgp[row_r * N + i];
            }
            if(tid == 0) {
                // trait_rfunction.h:50-56
                const double th = P.partial[interval] * (double)(k + 1);
                const double th2 = P.partial[interval] * (double)(P.nlod + 1 - (k + 1));
                tm.scal[0] = th; tm.scal[1] = 1.0 - th; tm.scal[2] = th2; tm.scal[3] = 1.0 - th2;
            }
            team_sync<T>(team);
            // trait_prob x recombination probability per child and (i, j) (trait_rfunction.cc:9-22)
            for(int q = tid; q < 4 * (N - F); q += T) {
                const int person = F + (q >> 2);
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
            // crossovers between the flanking markers (descent_graph.cc:212-242)
            for(int i = F + tid; i < N; i += T) {
                const uint32_t x = tm.dgl[i] ^ tm.dgr[i];
                ncross += (int)(x & 1u) + (sex_linked ? 0 : (int)((x >> 1) & 1u));
            }
        }
        else {
            for(int q = tid; q < 4 * (N - F); q += T) tm.tables[q] = trait_prob;
        }
        if(tid == 0) tm.red[0] = 0;
        team_sync<T>(team);
#pragma unroll
        for(int o = 16; o > 0; o >>= 1) ncross += __shfl_xor_sync(0xffffffffu, ncross, o);
        if(lane == 0 && ncross) atomicAdd(&tm.red[0], ncross);

        for(int lv = 0; lv < pr.n_flevels; ++lv) {
            const int b = pg.flevel_start[lv], e = pg.flevel_start[lv + 1];
            const uint32_t nq = pg.flevel_quads[lv];            // rows of this level
            if(nq * 4u <= 2u * T) {
                // latency mode: one (row, slot) per thread
                for(uint32_t q = tid; q < nq * 4u; q += T) {
                    const uint32_t row = q >> 2;
                    int lo = b, hi = e - 1;
                    while(lo < hi) { const int mid = (lo + hi + 1) >> 1; if(pg.lpf[mid] <= row) lo = mid; else hi = mid - 1; }
                    run_tile<true, DBG, 1>(L, pg, tm, pg.lops[lo], row - pg.lpf[lo], q & 3u);
                }
            }
            else {
                for(uint32_t q = tid; q < nq * (4u / SLK_TILE_NS); q += T) {
                    const uint32_t row = q / (4u / SLK_TILE_NS);
                    int lo = b, hi = e - 1;
                    while(lo < hi) { const int mid = (lo + hi + 1) >> 1; if(pg.lpf[mid] <= row) lo = mid; else hi = mid - 1; }
                    run_tile<true, DBG, SLK_TILE_NS>(L, pg, tm, pg.lops[lo], row - pg.lpf[lo], (q % (4u / SLK_TILE_NS)) * SLK_TILE_NS);
                }
            }
            team_sync<T>(team);
        }

        if(tid == 0) {
            const uint32_t* last_rec = pg.stream + pg.op_start[P.last_op];
            const double result = mat_load(mat_ref(tm, (int) last_rec[1]), 0);
            double prob = -DBL_MAX;
            if(result <= 0.0) raise_error(L.err, SLK_ERR_NONPOSITIVE_TRAIT, unit);      // peeler.cc:92-95
            else if(L.no_dg) prob = log(result);
            else {
                const int nmeioses = (sex_linked ? 1 : 2) * (N - F);
                const int nc = tm.red[0];
                const double recomb = (double) nc * P.log_theta[interval] + (double)(nmeioses - nc) * P.log_1mtheta[interval];
                prob = log(result) - recomb - P.marker_transmission;                     // peeler.cc:97-99
            }
            if(L.accumulate) {
                // LODscores::add = log_sum(prob, old) (lod_score.h:74-80, logarithms.cc:14-23)
                const double old = L.lod[unit];
                if(result > 0.0) L.lod[unit] = (old == -DBL_MAX) ? prob : log(exp(old - prob) + 1.0) + prob;
            }
            else {
                L.dump_result[u] = result;
                L.dump_prob[u] = prob;
            }
        }
        team_sync<T>(team);
    }
}

// ---- small utility kernels -------------------------------------------------------------

// int32[M][N][2] (descent_graph.h:35-37) <-> packed bytes
__global__ void slk_dg_pack_kernel(const int32_t* __restrict__ src, uint8_t* __restrict__ dst, size_t n) {
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if(i < n) {
        int2 v = ((const int2*) src)[i];
        dst[i] = (uint8_t)((v.x & 1) | ((v.y & 1) << 1));
    }
}

__global__ void slk_dg_unpack_kernel(const uint8_t* __restrict__ src, int32_t* __restrict__ dst, size_t n) {
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if(i < n) {
        uint8_t b = src[i];
        ((int2*) dst)[i] = make_int2(b & 1, (b >> 1) & 1);
    }
}

// run_gpu_lodscoreinit_kernel (cuda_lodscore.cu:505-509)
__global__ void slk_lod_init_kernel(double* lod, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) lod[i] = -DBL_MAX;
}

// run_gpu_lodscorenormalise_kernel (cuda_lodscore.cu:511-516), lod_score.h:86-88
__global__ void slk_lod_normalise_kernel(const double* lod, double* out, int n, double log_count, double trait_prob) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) out[i] = (lod[i] - log_count - trait_prob) / log(10.0);
}

__global__ void slk_philox_kernel(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
    uint32_t o[4];
    slk_philox4x32_10(c0, c1, c2, c3, k0, k1, o);
    for(int i = 0; i < 4; ++i) out[i] = o[i];
}

__global__ void slk_uniform_kernel(uint64_t seed, uint32_t chain, uint64_t iteration, uint32_t locus, uint32_t slot, double* out) {
    out[0] = slk_uniform(seed, chain, iteration, locus, slot);
}

// FP64 FMA peak: 8 independent chains per thread, enough warps to fill every SM
__global__ void slk_fp64_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for(int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

#endif
