// slk_msampler.cuh -- sm_100a kernels for the M-sampler (whole-chromosome Gibbs update of one
// meiosis indicator) and the descent-graph likelihood.
//
// Reference: MeiosisSampler::reset / step (meiosis_sampler.cc:17-203), FounderAlleleGraph4
// (founder_allele_graph4.cc:34-598), DescentGraph::get_likelihood (descent_graph.cc:150-265);
// the reference's own GPU attempt is run_gpu_msampler_*_kernel (cuda_common.h:240-244,
// cuda_msampler2.cu).  One step of the reference is
//     (1) per locus: likelihood of the founder allele graph with the meiosis flipped,
//     (2) a two-state forward pass along the chromosome, normalised at every locus,
//     (3) backward sampling of the indicator at every locus.
// (1) is independent per locus, (2) and (3) are recurrences along the chromosome.
//
// Execution model
//   slk_ms_likelihood_kernel   one THREAD per locus.  Every per-thread array lives in shared memory,
//                              interleaved by thread (element k of thread t sits in bank t), so the
//                              data-dependent indexing of the union-find is bank-conflict free:
//                                phase A  founder-allele labels of every person from the descent
//                                         graph row (one byte per person) in topological order;
//                                phase B  the labels of the typed people are compacted and the label
//                                         array is overlaid by the component tables;
//                                phase C  the reference's sequential graph walk, same operation order
//                                         (so the likelihood is bit-identical to the CPU's).
//                              Nothing is kept between steps: labels are recomputed from the graph row
//                              with the meiosis bit flipped, which is what FounderAlleleGraph4::flip +
//                              propagate_fa_update produce incrementally.
//   slk_ms_chain_kernel        one CTA.  (2) is a product of 2x2 non-negative matrices
//                              diag(raw_i) * T(theta_{i-1}): every thread multiplies the matrices of its
//                              chunk of loci, a block-wide scan (power-of-two rescaling, exact) gives each
//                              chunk its entry vector, and the chunk is then walked with the reference's
//                              own normalised recurrence.  (3) is a composition of maps {0,1} -> {0,1}
//                              (one per locus, fixed by that locus's Philox draw): same chunk / scan /
//                              walk structure.  Differs from the sequential CPU pass only by the rounding
//                              of the chunk entry vectors (<= 1e-12 relative, all terms non-negative).
#ifndef SLK_MSAMPLER_CUH
#define SLK_MSAMPLER_CUH

#include <stdint.h>
#include <float.h>

#include "slk_types.h"
#include "slk_philox.cuh"

#include <cooperative_groups.h>

#define SLK_SLOT_MSHUFFLE 0x7ffffff1u
#define SLK_SLOT_KIND     0x7ffffff2u
#define SLK_SLOT_MEIOSIS  0x40000000u


struct SlkMsLaunch {
    SlkMsPlan ms;
    int N, F, M, sex_linked;
    const int16_t* mother;
    const int16_t* father;
    const uint8_t* male;
    const double* theta;
    uint8_t* dgp;                // [M][N]
    double* cur;                 // [M] likelihood of the current graph at each locus
    double2* raw;                // [M] raw_matrix of the step: (meiosis = 0, meiosis = 1)
    uint8_t* bits;               // [M] the indicator's current value at each locus
    double* u;                   // [M] the step's indicator draws (Philox, keyed by locus)
    double* fb;                  // [M][2] forward matrix (parity dump; write-only for the kernels)
    int* err;
    int person, parent;          // the meiosis; person < 0: no flip, result goes to cur (reset)
    uint64_t seed, iteration;
    uint32_t chain, slot;
    int32_t* dump_edges;         // optional [M][2N]
    double* out;                 // dg likelihood: [0] = sum ln(lik), [1] = recombination term
    const double* log_theta;
    const double* log_1mtheta;
};

// ---- per-thread arrays in shared memory, interleaved so that thread t always hits bank t -------

struct MsLane {
    uint32_t base;               // shared-space byte address of the warp's slab + 4 * lane
    uint32_t base8;              // the same + 8 * lane (doubles)
    __device__ __forceinline__ uint32_t a8(uint32_t off, uint32_t k) const { return base + off + (((k & ~3u) << 5) | (k & 3u)); }
    __device__ __forceinline__ uint32_t a16(uint32_t off, uint32_t k) const { return base + off + (((k & ~1u) << 6) | ((k & 1u) << 1)); }
    __device__ __forceinline__ uint32_t a32(uint32_t off, uint32_t k) const { return base + off + (k << 7); }
    __device__ __forceinline__ uint32_t a64(uint32_t off, uint32_t k) const { return base8 + off + (k << 8); }
};

// plain (non-volatile) accessors: the compiler may reorder and batch them like ordinary loads and
// stores; the "memory" clobber on the stores keeps a store ahead of later loads of the same array
__device__ __forceinline__ uint32_t ms_ld8(uint32_t a)  { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t ms_ld16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ double   ms_ld64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void ms_st8(uint32_t a, uint32_t v)  { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void ms_st16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void ms_st32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void ms_st64(uint32_t a, double v)   { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory"); }

// Host and device agree on the carve-up through this function.  Per thread, in bytes:
//   keep     te (labels of typed people, 2*nt entries of E bytes)
//   overlay  phase A: labels (2N entries of E bytes) | graph row (N bytes)
//            phase C: prob0, prob1 (G doubles each) | fa (2F x u16) | grp (G x u16)
// E = 1 if 2F <= 256 else 2; G = min(nt, 2F) components at most (each one is created by a typed
// person and claims at least one founder allele).
struct MsLayout {
    int wide;                    // E == 2
    int G;
    uint32_t te, ov;             // region offsets for one thread (multiply by 32 lanes)
    uint32_t lab, row;           // phase A inside the overlay
    uint32_t fa, grp, prob0, prob1;     // phase C inside the overlay
    uint32_t per_thread;         // bytes per thread
    uint32_t cta_tables;         // bytes of CTA-shared tables (seq, mother, father, typed list, x-male flags)
};

#if defined(__CUDACC__)
__host__ __device__
#endif
static inline MsLayout slk_ms_layout(int N, int F, int nt) {
    MsLayout L;
    L.wide = (2 * F > 256) ? 1 : 0;
    L.G = nt < 2 * F ? nt : 2 * F;
    if(L.G < 1) L.G = 1;
    const uint32_t E = L.wide ? 2u : 1u;
#define r4(x) ((((uint32_t)(x)) + 3u) & ~3u)
#define r8(x) ((((uint32_t)(x)) + 7u) & ~7u)
    L.te = 0;
    L.ov = r8(L.te + r4(2u * nt * E));
    L.lab = 0;
    L.row = L.lab + r4(2u * N * E);
    const uint32_t a_bytes = L.row + r4((uint32_t) N);
    L.prob0 = 0;                                          // doubles first: keeps them 8-byte aligned
    L.prob1 = L.prob0 + 8u * L.G;
    L.fa = L.prob1 + 8u * L.G;
    L.grp = L.fa + r4(4u * F);
    const uint32_t c_bytes = L.grp + r4(2u * L.G);
    L.per_thread = L.ov + r8(a_bytes > c_bytes ? a_bytes : c_bytes);
    L.cta_tables = r8((uint32_t)(2 * (N - F) + 4 * N + 2 * nt + nt + 8));
#undef r4
#undef r8
    return L;
}

// ---- (1) founder allele graph likelihood, one thread per locus ---------------------------------
//
// The reference relabels every founder allele of the absorbed component on each merge
// (combine_components, founder_allele_graph4.cc:504-546: a loop over all 2F alleles).  Here the
// components form a disjoint-set forest instead: a founder allele remembers the component it first
// joined and its two candidate alleles in that component's frame; a component that is merged away
// points at its absorber with a parity bit ("my two assignments are swapped relative to yours").
// find() is a couple of hops with path compression.  Component ids are creation order and the
// absorber is always the maternal allele's component, as in the reference, so every product is
// formed from the same factors in the same order: the likelihood is bit-identical.
//
// grp entry (u16): bits 0..11 parent component, bit 12 parity to parent, bits 13..14 state
// (0 unfixed, 1 fixed to assignment 0, 2 fixed to assignment 1), bit 15 merged away.
// fa entry (u16):  bits 0..11 component + 1 (0 = none), bit 14 / 15 = candidate allele under
// assignment 0 / 1 is HOMOZ_B (in the frame of that component).

#define SLK_MS_MAXGROUPS 4095

template<bool WIDE>
__global__ void __launch_bounds__(32)
slk_ms_likelihood_kernel(const SlkMsLaunch L) {
    extern __shared__ __align__(16) unsigned char ms_smem[];
    const int N = L.N, F = L.F, M = L.M, nt = L.ms.n_typed;
    const MsLayout lay = slk_ms_layout(N, F, nt);
    const uint32_t T = blockDim.x, t = threadIdx.x;

    // CTA-shared tables
    uint16_t* s_seq = (uint16_t*) ms_smem;                 // [N-F]
    int16_t* s_mo = (int16_t*)(s_seq + (N - F));          // [N]
    int16_t* s_fa = s_mo + N;                             // [N]
    uint16_t* s_typed = (uint16_t*)(s_fa + N);            // [nt]
    uint8_t* s_auto = (uint8_t*)(s_typed + nt);           // [nt] x-linked male: maternal allele only
    for(int i = t; i < N - F; i += T) s_seq[i] = L.ms.seq[i];
    for(int i = t; i < N; i += T) { s_mo[i] = L.mother[i]; s_fa[i] = L.father[i]; }
    for(int i = t; i < nt; i += T) {
        int p = L.ms.typed[i];
        s_typed[i] = (uint16_t) p;
        s_auto[i] = (L.sex_linked && L.male[p]) ? 1 : 0;
    }
    // every warp owns a contiguous slab, interleaved by lane.  The phase A and phase C tables overlay
    // each other ACROSS the lanes of a warp (a double of lane t covers the 32-bit columns 2t and 2t+1),
    // so the phases are separated by __syncwarp() and the locus loop is uniform per warp.
    MsLane ln;
    {
        const uint32_t slab = (uint32_t) __cvta_generic_to_shared(ms_smem + lay.cta_tables) + (t >> 5) * 32u * lay.per_thread;
        ln.base = slab + ((t & 31u) << 2);
        ln.base8 = slab + ((t & 31u) << 3);
    }
    const uint32_t o_te = lay.te * 32u, o_ov = lay.ov * 32u;
    const uint32_t o_lab = o_ov + lay.lab * 32u, o_row = o_ov + lay.row * 32u;
    const uint32_t o_fa = o_ov + lay.fa * 32u, o_grp = o_ov + lay.grp * 32u;
    const uint32_t o_p0 = o_ov + lay.prob0 * 32u, o_p1 = o_ov + lay.prob1 * 32u;
    __syncthreads();

    // loci are dealt out evenly: CTA b owns [b*per, (b+1)*per), thread t the t-th, t+T-th, ... of them
    const int per = (M + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per;
    const int hi = min(M, lo + per);

#define LAB_LD(k)     (WIDE ? ms_ld16(ln.a16(o_lab, (k))) : ms_ld8(ln.a8(o_lab, (k))))
#define LAB_ST(k, v)  do { if(WIDE) ms_st16(ln.a16(o_lab, (k)), (v)); else ms_st8(ln.a8(o_lab, (k)), (v)); } while(0)
#define TE_LD(k)      (WIDE ? ms_ld16(ln.a16(o_te, (k))) : ms_ld8(ln.a8(o_te, (k))))
#define TE_ST(k, v)   do { if(WIDE) ms_st16(ln.a16(o_te, (k)), (v)); else ms_st8(ln.a8(o_te, (k)), (v)); } while(0)
#define FA_LD(k)      ms_ld16(ln.a16(o_fa, (k)))
#define FA_ST(k, v)   ms_st16(ln.a16(o_fa, (k)), (v))
#define GRP_LD(k)     ms_ld16(ln.a16(o_grp, (k)))
#define GRP_ST(k, v)  ms_st16(ln.a16(o_grp, (k)), (v))
#define P_LD(w, k)    ms_ld64(ln.a64((w) ? o_p1 : o_p0, (k)))
#define P_ST(w, k, v) ms_st64(ln.a64((w) ? o_p1 : o_p0, (k)), (v))
// founder labels are implicit: person p < F carries alleles 2p, 2p + 1
#define LABEL_OF(p, b) ((p) < F ? (uint32_t)(2 * (p) + (b)) : LAB_LD(2 * (p) + (b)))

    for(int l0 = lo + (int)(t & ~31u); l0 < hi; l0 += (int) T) {
        const int l = l0 + (int)(t & 31u);
        const bool live = l < hi;
        __syncwarp();                                  // the previous locus's phase C tables are dead
        // ---- phase A: labels -------------------------------------------------------------------
        uint32_t mybit = 0;
        if(live) {
            const uint8_t* row = L.dgp + (size_t) l * N;
            if(((N | (int)(size_t) L.dgp) & 3) == 0) {
                const uint32_t* row4 = (const uint32_t*) row;
                for(int i = F >> 2; i < (N >> 2); ++i) ms_st32(ln.a32(o_row, i), __ldg(row4 + i));
            }
            else {
                for(int i = F; i < N; ++i) ms_st8(ln.a8(o_row, i), row[i]);
            }
            for(int k = 0; k < N - F; ++k) {
                const int i = s_seq[k];
                uint32_t b = ms_ld8(ln.a8(o_row, i));
                if(i == L.person) { mybit = (b >> L.parent) & 1u; b ^= (1u << L.parent); }
                const int mo = s_mo[i], fa = s_fa[i];
                const uint32_t m = LABEL_OF(mo, b & 1u);
                const uint32_t f = LABEL_OF(fa, (b >> 1) & 1u);
                LAB_ST(2 * i, m);
                LAB_ST(2 * i + 1, f);
            }
            if(L.dump_edges) {
                int32_t* e = L.dump_edges + (size_t) l * 2 * N;
                for(int i = 0; i < 2 * N; ++i) e[i] = (int32_t) LABEL_OF(i >> 1, i & 1);
            }
        }
        // ---- phase B: keep the typed people's labels, hand the overlay to the component tables ----
        if(live) {
            for(int k = 0; k < nt; ++k) {
                const int i = s_typed[k];
                const uint32_t a = LABEL_OF(i, 0), b = LABEL_OF(i, 1);
                TE_ST(2 * k, a);
                TE_ST(2 * k + 1, b);
            }
        }
        __syncwarp();                                  // every lane is done with its labels
        if(!live) continue;
        for(int i = 0; i < F; ++i) ms_st32(ln.a32(o_fa, i), 0u);       // two u16 entries per word

        // ---- phase C: founder_allele_graph4.cc:34-424 -------------------------------------------
        const double minor = L.ms.minor[l], major = 1.0 - minor;
        int ngroups = 0;
        bool dead = false;

        // root of component g: returns the root id, its entry in `re`, the parity of g's frame to the root's
        auto find = [&](uint32_t g, uint32_t& re, uint32_t& par) -> uint32_t {
            const uint32_t g0 = g;
            uint32_t e = GRP_LD(g);
            par = 0;
            int hops = 0;
            while((e & 0xfffu) != g) {
                par ^= (e >> 12) & 1u;
                g = e & 0xfffu;
                e = GRP_LD(g);
                ++hops;
            }
            if(hops > 1) GRP_ST(g0, g | (par << 12) | 0x8000u);    // path compression
            re = e;
            return g;
        };
        // state field -> the reference's group_fixed (-1 unfixed, 0, 1)
        auto fixed_of = [](uint32_t e) -> int { return (int)((e >> 13) & 3u) - 1; };
        auto set_fixed = [&](uint32_t root, int f) { GRP_ST(root, root | ((uint32_t)(f + 1) << 13)); };

        // observed genotypes come straight from global memory ([nt][M]: coalesced over the lanes), one
        // person ahead of their use
        const uint8_t* obs = L.ms.obsT + l;
        uint32_t g_next = nt > 0 ? __ldg(obs) : 0u;
        for(int k = 0; k < nt && !dead; ++k) {
            const uint32_t g = g_next;
            if(k + 1 < nt) g_next = __ldg(obs + (size_t)(k + 1) * M);
            if(g == SLK_UNTYPED) continue;
            const uint32_t mat = TE_LD(2 * k), pat = TE_LD(2 * k + 1);
            const uint32_t gB = (g == SLK_HOMOZ_B) ? 1u : 0u;
            const bool het = g == SLK_HETERO;
            const uint32_t em = FA_LD(mat);

            if(mat == pat || s_auto[k]) {
                if(het) { dead = true; break; }
                if(em & 0xfffu) {
                    uint32_t re, par;
                    const uint32_t g1 = find((em & 0xfffu) - 1u, re, par);
                    const int f1 = fixed_of(re);
                    const uint32_t a0 = (em >> (14 + par)) & 1u, a1 = (em >> (15 - par)) & 1u;   // root frame
                    if(f1 != -1) {
                        if((f1 ? a1 : a0) != gB) { dead = true; break; }
                    }
                    else {
                        if(a0 == gB)      { set_fixed(g1, 0); P_ST(1, g1, 0.0); }
                        else if(a1 == gB) { set_fixed(g1, 1); P_ST(0, g1, 0.0); }
                        else { dead = true; break; }
                    }
                }
                else {
                    FA_ST(mat, (uint32_t)(ngroups + 1) | (gB << 14));
                    set_fixed(ngroups, 0);
                    P_ST(0, ngroups, gB ? minor : major);
                    P_ST(1, ngroups, 0.0);
                    ++ngroups;
                }
                continue;
            }

            const uint32_t ep = FA_LD(pat);
            // legal(g, a1, a2) and get_other_allele(g, a1) on one-bit alleles (founder_allele_graph4.cc:427-455)
#define LEGAL(a1, a2) (het ? ((a1) != (a2)) : ((a1) == gB && (a2) == gB))
            if((em & 0xfffu) && (ep & 0xfffu)) {
                uint32_t re1, par1, re2, par2;
                const uint32_t g1 = find((em & 0xfffu) - 1u, re1, par1);
                const uint32_t g2 = find((ep & 0xfffu) - 1u, re2, par2);
                const uint32_t m0 = (em >> (14 + par1)) & 1u, m1 = (em >> (15 - par1)) & 1u;
                const uint32_t q0 = (ep >> (14 + par2)) & 1u, q1 = (ep >> (15 - par2)) & 1u;
                int f1 = fixed_of(re1);
                if(g1 == g2) {
                    if(f1 != -1) {
                        if(!LEGAL(f1 ? m1 : m0, f1 ? q1 : q0)) { dead = true; break; }
                    }
                    else {
                        const bool l0 = LEGAL(m0, q0), l1 = LEGAL(m1, q1);
                        if(l0) { if(!l1) { set_fixed(g1, 0); P_ST(1, g1, 0.0); } }
                        else {
                            if(l1) { set_fixed(g1, 1); P_ST(0, g1, 0.0); }
                            else { dead = true; break; }
                        }
                    }
                    continue;
                }
                int f2 = fixed_of(re2);
                if(f1 != -1) {
                    const uint32_t am = f1 ? m1 : m0;
                    if(f2 != -1) {
                        if(!LEGAL(am, f2 ? q1 : q0)) { dead = true; break; }
                    }
                    else {
                        if(LEGAL(am, q0)) f2 = 0;
                        else if(LEGAL(am, q1)) f2 = 1;
                        else { dead = true; break; }
                    }
                }
                else if(f2 != -1) {
                    const uint32_t ap = f2 ? q1 : q0;
                    if(LEGAL(m0, ap)) f1 = 0;
                    else if(LEGAL(m1, ap)) f1 = 1;
                    else { dead = true; break; }
                }
                else {
                    const bool l0 = LEGAL(m0, q0), l1 = LEGAL(m1, q0), l2 = LEGAL(m0, q1), l3 = LEGAL(m1, q1);
                    if(!(l0 || l1 || l2 || l3)) { dead = true; break; }
                    if(l0 && !(l1 || l2 || l3))      { f1 = f2 = 0; }
                    else if(l1 && !(l0 || l2 || l3)) { f1 = 1; f2 = 0; }
                    else if(l2 && !(l0 || l1 || l3)) { f1 = 0; f2 = 1; }
                    else if(l3 && !(l0 || l1 || l2)) { f1 = f2 = 1; }
                    else if(l0 && l3 && !(l1 || l2)) { f1 = f2 = -1; }
                    else if(l1 && l2 && !(l0 || l3)) { f1 = f2 = -2; }
                    else                             { f1 = f2 = -1; }
                }
                bool flip;
                double a0 = P_LD(0, g1), a1 = P_LD(1, g1), b0 = P_LD(0, g2), b1 = P_LD(1, g2);
                if(f1 != f2) {
                    set_fixed(g1, f1);
                    if(f1) a0 = 0.0; else a1 = 0.0;            // prob[1-fixed1][group1] = 0
                    if(f2) b0 = 0.0; else b1 = 0.0;            // prob[1-fixed2][group2] = 0
                    flip = true;
                }
                else if(f1 == -2) {
                    set_fixed(g1, -1);
                    flip = true;
                }
                else {
                    set_fixed(g1, f1);
                    if(f1 != -1) {
                        if(f1) { a0 = 0.0; b0 = 0.0; } else { a1 = 0.0; b1 = 0.0; }
                    }
                    flip = false;
                }
                // combine_components(group1, group2, flip) (:504-546): group2 now points at group1
                GRP_ST(g2, g1 | ((flip ? 1u : 0u) << 12) | 0x8000u);
                if(flip) { P_ST(0, g1, a0 * b1); P_ST(1, g1, a1 * b0); }
                else     { P_ST(0, g1, a0 * b0); P_ST(1, g1, a1 * b1); }
                continue;
            }
#undef LEGAL
            if((em & 0xfffu) || (ep & 0xfffu)) {
                // one allele already in a component, the other joins it (:286-384)
                const bool mat_in = (em & 0xfffu) != 0;
                const uint32_t eo = mat_in ? em : ep;
                uint32_t re, par;
                const uint32_t grp = find((eo & 0xfffu) - 1u, re, par);
                const uint32_t fa_new = mat_in ? pat : mat;
                const uint32_t o0 = (eo >> (14 + par)) & 1u, o1 = (eo >> (15 - par)) & 1u;
                const int fx = fixed_of(re);
                uint32_t v0 = 0, v1 = 0;
                // other allele under assignment w: valid iff het or o_w == gB; value = het ? !o_w : gB
                if(fx != -1) {
                    const uint32_t o = fx ? o1 : o0;
                    if(!het && o != gB) { dead = true; break; }
                    const uint32_t v = het ? (o ^ 1u) : gB;
                    if(fx) v1 = v; else v0 = v;
                    P_ST(fx, grp, P_LD(fx, grp) * (v ? minor : major));
                }
                else {
                    const bool ok0 = het || o0 == gB, ok1 = het || o1 == gB;
                    v0 = het ? (o0 ^ 1u) : gB; v1 = het ? (o1 ^ 1u) : gB;
                    if(ok0) {
                        if(ok1) {
                            P_ST(0, grp, P_LD(0, grp) * (v0 ? minor : major));
                            P_ST(1, grp, P_LD(1, grp) * (v1 ? minor : major));
                        }
                        else {
                            P_ST(0, grp, P_LD(0, grp) * (v0 ? minor : major));
                            P_ST(1, grp, 0.0);
                            set_fixed(grp, 0);
                        }
                    }
                    else {
                        if(ok1) {
                            P_ST(1, grp, P_LD(1, grp) * (v1 ? minor : major));
                            P_ST(0, grp, 0.0);
                            set_fixed(grp, 1);
                        }
                        else { dead = true; break; }
                    }
                }
                FA_ST(fa_new, (grp + 1u) | (v0 << 14) | (v1 << 15));
                continue;
            }
            // neither allele seen before: a new component (:386-409)
            if(het) {
                FA_ST(mat, (uint32_t)(ngroups + 1) | (0u << 14) | (1u << 15));
                FA_ST(pat, (uint32_t)(ngroups + 1) | (1u << 14) | (0u << 15));
                const double pr = major * minor;
                P_ST(0, ngroups, pr);
                P_ST(1, ngroups, pr);
                set_fixed(ngroups, -1);
            }
            else {
                FA_ST(mat, (uint32_t)(ngroups + 1) | (gB << 14));
                FA_ST(pat, (uint32_t)(ngroups + 1) | (gB << 14));
                const double fq = gB ? minor : major;
                P_ST(0, ngroups, fq * fq);
                P_ST(1, ngroups, 0.0);
                set_fixed(ngroups, 0);
            }
            ++ngroups;
        }

        double ret = 0.0;
        if(!dead) {
            ret = 1.0;
            for(int i = 0; i < ngroups; ++i) {
                const uint32_t e = GRP_LD(i);
                if(e & 0x8000u) continue;                  // merged away (group_active false)
                const int fx = fixed_of(e);
                if(fx != -1) ret *= P_LD(fx, i);
                else ret *= (P_LD(0, i) + P_LD(1, i));
            }
        }
        if(L.person < 0) L.cur[l] = ret;
        else {
            const double c = L.cur[l];
            L.raw[l] = mybit ? make_double2(ret, c) : make_double2(c, ret);   // meiosis_sampler.cc:126-127
            L.bits[l] = (uint8_t) mybit;
            L.u[l] = slk_uniform(L.seed, L.chain, L.iteration, (uint32_t) l, L.slot);
        }
    }
#undef LAB_LD
#undef LAB_ST
#undef TE_LD
#undef TE_ST
#undef FA_LD
#undef FA_ST
#undef GRP_LD
#undef GRP_ST
#undef P_LD
#undef P_ST
#undef LABEL_OF
}

// ---- (2) + (3): forward pass and backward sampling along the chromosome, one CTA ------------------

struct Mat2 { double a, b, c, d; };     // [[a b][c d]]

__device__ __forceinline__ Mat2 mat2_mul_rescale(const Mat2& x, const Mat2& y) {
    // x * y, all entries non-negative, then scaled by a power of two so the largest is in [1, 2)
    Mat2 r;
    r.a = x.a * y.a + x.b * y.c;
    r.b = x.a * y.b + x.b * y.d;
    r.c = x.c * y.a + x.d * y.c;
    r.d = x.c * y.b + x.d * y.d;
    const double m = fmax(fmax(r.a, r.b), fmax(r.c, r.d));
    if(m > 0.0 && m < DBL_MAX) {
        const int e = ((__double2hiint(m) >> 20) & 0x7ff) - 1023;
        if(e != 0 && e > -1000 && e < 1000) {
            const double s = __hiloint2double((1023 - e) << 20, 0);
            r.a *= s; r.b *= s; r.c *= s; r.d *= s;
        }
    }
    return r;
}

// meiosis_sampler.cc:193-203
__device__ __forceinline__ int ms_pick(double w0, double w1, double u) {
    if(w0 == 0.0) return 1;
    if(w1 == 0.0) return 0;
    return (u < (w0 / (w0 + w1))) ? 0 : 1;
}

__device__ __forceinline__ Mat2 mat2_shfl_up(const Mat2& v, int d) {
    Mat2 r;
    r.a = __shfl_up_sync(0xffffffffu, v.a, d); r.b = __shfl_up_sync(0xffffffffu, v.b, d);
    r.c = __shfl_up_sync(0xffffffffu, v.c, d); r.d = __shfl_up_sync(0xffffffffu, v.d, d);
    return r;
}

// One thread-block CLUSTER of SLK_MS_CLUSTER CTAs (the chromosome is cut into contiguous chunks, one
// per thread of the cluster); the two scans cross CTAs through distributed shared memory.  Dynamic
// shared memory of a CTA: the raw pairs of its own loci (16 B each, staged with coalesced loads) and
// one map byte per locus.
#define SLK_MS_CLUSTER 8
#define SLK_MS_CHAIN_THREADS 512

__global__ void __cluster_dims__(SLK_MS_CLUSTER, 1, 1) __launch_bounds__(SLK_MS_CHAIN_THREADS)
slk_ms_chain_kernel(const SlkMsLaunch L) {
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) unsigned char ch_smem[];
    __shared__ Mat2 s_wmat[32];
    __shared__ Mat2 s_cmat[SLK_MS_CLUSTER];        // the product of each CTA of the cluster
    __shared__ uint8_t s_wmap[32];
    __shared__ uint8_t s_cmap[SLK_MS_CLUSTER];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int) cluster.block_rank();
    const int M = L.M, N = L.N, T = blockDim.x, t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5, nwarps = T >> 5;
    const int C = (M + SLK_MS_CLUSTER * T - 1) / (SLK_MS_CLUSTER * T);      // loci per thread
    const int cta_lo = min(M, rank * T * C), cta_hi = min(M, cta_lo + T * C);
    const int lo = min(M, cta_lo + t * C), hi = min(M, lo + C);
    double2* s_raw = (double2*) ch_smem;                                    // [T*C]
    uint8_t* s_map = ch_smem + (size_t) T * C * sizeof(double2);            // [T*C]
    for(int i = cta_lo + t; i < cta_hi; i += T) s_raw[i - cta_lo] = L.raw[i];
    __syncthreads();

    // ---- forward: product of the chunk's matrices A_i = diag(raw_i) * [[1-th, th], [th, 1-th]] ----
    Mat2 acc; acc.a = 1.0; acc.b = 0.0; acc.c = 0.0; acc.d = 1.0;
    int bad = -1;
    for(int i = lo; i < hi; ++i) {
        const double2 r = s_raw[i - cta_lo];
        if(r.x == 0.0 && r.y == 0.0) bad = i;
        if(i == 0) continue;                       // locus 0 enters as the start vector
        const double th = __ldg(L.theta + i - 1), ith = 1.0 - th;
        Mat2 A; A.a = r.x * ith; A.b = r.x * th; A.c = r.y * th; A.d = r.y * ith;
        acc = mat2_mul_rescale(A, acc);
    }
    if(bad >= 0 && atomicCAS(&L.err[0], 0, SLK_ERR_ILLEGAL_GRAPH) == 0) L.err[1] = bad;
    // inclusive scan over the threads of the CTA, later chunks multiply from the left
    for(int d = 1; d < 32; d <<= 1) {
        const Mat2 o = mat2_shfl_up(acc, d);
        if(lane >= d) acc = mat2_mul_rescale(acc, o);
    }
    if(lane == 31) s_wmat[warp] = acc;
    __syncthreads();
    if(warp == 0) {
        Mat2 w; w.a = 1.0; w.b = 0.0; w.c = 0.0; w.d = 1.0;
        if(lane < nwarps) w = s_wmat[lane];
        for(int d = 1; d < 32; d <<= 1) {
            const Mat2 o = mat2_shfl_up(w, d);
            if(lane >= d) w = mat2_mul_rescale(w, o);
        }
        s_wmat[lane] = w;
    }
    __syncthreads();
    if(t < SLK_MS_CLUSTER) {
        Mat2* remote = cluster.map_shared_rank(s_cmat, t);
        remote[rank] = s_wmat[nwarps - 1];
    }
    cluster.sync();

    // meiosis_sampler.cc:155-187 for one locus, both possible values of the next locus's indicator:
    // bit 0 = image of 0, bit 1 = image of 1
    auto map_code = [&](int i, double f0, double f1) -> uint32_t {
        const double u = L.u[i];
        if(i == M - 1) { const int x = ms_pick(f0, f1, u); return (uint32_t)(x | (x << 1)); }
        const double th = __ldg(L.theta + i), ith = 1.0 - th;
        const int x0 = ms_pick(f0 * ith, f1 * th, u);     // next = 0: j != next ? theta : 1 - theta
        const int x1 = ms_pick(f0 * th, f1 * ith, u);     // next = 1
        return (uint32_t)(x0 | (x1 << 1));
    };
    auto apply = [](uint32_t f, uint32_t x) -> uint32_t { return (f >> x) & 1u; };
    auto compose = [&](uint32_t f, uint32_t g) -> uint32_t { return apply(f, apply(g, 0)) | (apply(f, apply(g, 1)) << 1); };

    uint32_t gmap = 2u;                              // identity
    {
        // exclusive prefix of this thread = (lane before) x (warps before) x (CTAs before)
        Mat2 q = mat2_shfl_up(acc, 1);
        if(lane == 0) { q.a = 1.0; q.b = 0.0; q.c = 0.0; q.d = 1.0; }
        if(warp > 0) q = mat2_mul_rescale(q, s_wmat[warp - 1]);
        for(int c = rank - 1; c >= 0; --c) q = mat2_mul_rescale(q, s_cmat[c]);
        if(lo < hi) {
            double2 r = L.raw[0];
            const double tot0 = r.x + r.y;
            double v0 = r.x / tot0, v1 = r.y / tot0;   // fb_matrix[0..1] (meiosis_sampler.cc:134-137)
            int i = lo;
            if(lo == 0) {
                L.fb[0] = v0; L.fb[1] = v1;
                s_map[0] = (uint8_t) map_code(0, v0, v1);
                i = 1;
            }
            else {
                const double w0 = q.a * v0 + q.b * v1, w1 = q.c * v0 + q.d * v1, tot = w0 + w1;
                v0 = w0 / tot; v1 = w1 / tot;
            }
            for(; i < hi; ++i) {                       // the reference's recurrence (:140-153), same operation order
                r = s_raw[i - cta_lo];
                const double th = __ldg(L.theta + i - 1), ith = 1.0 - th;
                double f0 = r.x * ((v1 * th) + (v0 * ith));
                double f1 = r.y * ((v0 * th) + (v1 * ith));
                const double tot = f0 + f1;
                f0 /= tot; f1 /= tot;
                L.fb[2 * i] = f0; L.fb[2 * i + 1] = f1;
                s_map[i - cta_lo] = (uint8_t) map_code(i, f0, f1);
                v0 = f0; v1 = f1;
            }
            for(i = hi - 1; i >= lo; --i) gmap = compose(s_map[i - cta_lo], gmap);
        }
    }

    // ---- backward: inclusive suffix scan of the maps, S_t = G_t o G_{t+1} o ... -----------------------
    for(int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_down_sync(0xffffffffu, gmap, d);
        if(lane + d < 32) gmap = compose(gmap, o);
    }
    if(lane == 0) s_wmap[warp] = (uint8_t) gmap;
    __syncthreads();
    if(warp == 0) {
        uint32_t w = lane < nwarps ? s_wmap[lane] : 2u;
        for(int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_down_sync(0xffffffffu, w, d);
            if(lane + d < 32) w = compose(w, o);
        }
        s_wmap[lane] = (uint8_t) w;
    }
    __syncthreads();
    if(t < SLK_MS_CLUSTER) {
        uint8_t* remote = cluster.map_shared_rank(s_cmap, t);
        remote[rank] = s_wmap[0];                      // the whole CTA's composition
    }
    cluster.sync();
    {
        // the map from "right of everything" to the value entering this chunk from the right; the last
        // locus's map is constant, so the composition is too and may be evaluated at 0
        uint32_t in = __shfl_down_sync(0xffffffffu, gmap, 1);
        if(lane == 31) in = 2u;
        if(warp + 1 < nwarps) in = compose(in, s_wmap[warp + 1]);
        for(int c = rank + 1; c < SLK_MS_CLUSTER; ++c) in = compose(in, s_cmap[c]);
        if(lo < hi) {
            uint32_t x = apply(in, 0);
            const uint32_t pbit = 1u << L.parent;
            for(int i = hi - 1; i >= lo; --i) {
                x = apply(s_map[i - cta_lo], x);
                const uint32_t bit = L.bits[i] & 1u;
                if(bit != x) {
                    uint8_t* cell = L.dgp + (size_t) i * N + L.person;
                    *cell = (uint8_t)(*cell ^ pbit);
                    const double2 r = s_raw[i - cta_lo];
                    L.cur[i] = x ? r.y : r.x;            // carried likelihood of the (new) current graph
                }
            }
        }
    }
    cluster.sync();                                    // no CTA may exit while its shared memory can still be read
}

// ---- descent-graph likelihood (descent_graph.cc:150-265) ------------------------------------------

// out[0] = sum over loci of ln(cur[l]) (or -DBL_MAX if any is 0), out[1] = sum over intervals of the
// recombination term; one CTA, fixed-order tree reduction so the result does not depend on timing
__global__ void __launch_bounds__(1024)
slk_ms_dg_likelihood_kernel(const SlkMsLaunch L) {
    __shared__ double s_a[1024], s_b[1024];
    __shared__ int s_zero;
    const int M = L.M, N = L.N, F = L.F, T = blockDim.x, t = threadIdx.x;
    if(t == 0) s_zero = 0;
    __syncthreads();
    double a = 0.0, b = 0.0;
    const uint32_t mask = L.sex_linked ? 1u : 3u;
    for(int l = t; l < M; l += T) {
        const double v = L.cur[l];
        if(v == 0.0) s_zero = 1; else a += log(v);
        if(l + 1 < M) {
            const uint8_t* r0 = L.dgp + (size_t) l * N;
            const uint8_t* r1 = r0 + N;
            int cross = 0;
            for(int i = F; i < N; ++i) cross += __popc((uint32_t)(r0[i] ^ r1[i]) & mask);
            const int total = (N - F) * (L.sex_linked ? 1 : 2);
            b += cross * L.log_theta[l] + (total - cross) * L.log_1mtheta[l];
        }
    }
    s_a[t] = a; s_b[t] = b;
    __syncthreads();
    for(int d = T >> 1; d > 0; d >>= 1) {
        if(t < d) { s_a[t] += s_a[t + d]; s_b[t] += s_b[t + d]; }
        __syncthreads();
    }
    if(t == 0) {
        L.out[0] = s_zero ? -DBL_MAX : s_a[0];
        L.out[1] = s_b[0];
    }
}

#endif
