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
// (1) is independent per locus, (2) and (3) are recurrences along the chromosome; consecutive
// steps depend on each other only through the one indicator that (3) samples.
//
// Execution model
//   slk_ms_likelihood_kernel   one THREAD per (locus, hypothesis), one warp per CTA.  Every per-thread
//                              array lives in shared memory, interleaved by lane (element k of lane t
//                              sits in bank t), so the data-dependent indexing of the union-find is
//                              bank-conflict free:
//                                phase A  founder-allele labels of every person from the descent
//                                         graph row (one byte per person) in topological order, with the
//                                         hypothesis's indicator bits flipped -- what
//                                         FounderAlleleGraph4::flip + propagate_fa_update produce;
//                                phase B  the labels of the typed people are compacted and the label
//                                         array is overlaid by the component tables;
//                                phase C  the reference's walk over the typed people, as a disjoint-set
//                                         forest over the founder alleles with parity bits instead of its
//                                         relabelling loops (ms_walk).
//                              Every factor of the likelihood is the major or the minor allele frequency
//                              of the locus, so a component carries two small counts, not two doubles,
//                              and the kernel returns ln L = A ln(major) + B ln(minor) + sum over unfixed
//                              components of ln(1 + (minor/major)^d): exact integer bookkeeping, one
//                              rounding step at the end (<= 1e-13 relative on L; north_star allows 1e-9).
//                              A launch evaluates up to three hypotheses: "meiosis j flipped", and for the
//                              NEXT meiosis of the sweep "j+1 flipped" / "j and j+1 flipped", i.e. both
//                              outcomes of step j -- so two steps cost one likelihood latency.
//   slk_ms_step_kernel         the same likelihood incrementally (only the labels a hypothesis can change are
//                              re-derived), used for every step of a sweep; launched with programmatic stream
//                              serialisation.  Two of these launches are in flight: a launch starts once the chain
//                              kernel before the previous one has completed, walks the typed people the last two
//                              pairs cannot affect, waits for its predecessor and walks the rest.  Every fourth
//                              launch also saves the forest over the people no meiosis of the next launches can
//                              affect (a snapshot); those launches start their walks from it.
//   slk_ms_chain_kernel        one 8-CTA cluster, one or two steps.  (2) is a product of 2x2 non-negative
//                              matrices diag(raw_i) * T(theta_{i-1}): every thread multiplies the matrices
//                              of its chunk of loci, a scan over threads / warps / CTAs (distributed
//                              shared memory; power-of-two rescaling, exact) gives each chunk its entry
//                              vector, and the chunk is walked with the reference's own normalised
//                              recurrence.  (3) is a composition of maps {0,1} -> {0,1} (one per locus,
//                              fixed by that locus's Philox draw): same chunk / scan / walk structure.
#ifndef SLK_MSAMPLER_CUH
#define SLK_MSAMPLER_CUH

#include <stdint.h>
#include <float.h>
#include <math.h>

#include "slk_types.h"
#include "slk_philox.cuh"

#include <cooperative_groups.h>

#define SLK_SLOT_MSHUFFLE 0x7ffffff1u
#define SLK_SLOT_KIND     0x7ffffff2u
#define SLK_SLOT_MEIOSIS  0x40000000u

#define SLK_MS_MAXSETS 3
#define SLK_MS_REC_WORDS 448
#define SLK_MS_MAXGROUPS 4095

struct SlkMsLaunch {
    SlkMsPlan ms;
    int N, F, M, sex_linked;
    const int16_t* mother;
    const int16_t* father;
    const uint8_t* male;
    const double* theta;
    uint8_t* dgp;                // [M][N]
    double* lncur;               // [M] ln likelihood of the current graph at each locus (-inf: impossible)
    double* lnl;                 // [3][M] ln likelihood under each hypothesis of the launch
    uint8_t* bits;               // [2][M] current value of the indicator of step 0 / step 1 at each locus
    double* fb;                  // [M][2] forward matrix of step 0 (parity dump; write-only for the kernels)
    void* te;                    // [nt][M] label pairs of the typed people in the current graph (u16, wide: u32)
    uint32_t* stale;             // [M][W] te slots (2k + parent) whose entry is out of date at that locus
    int* err;
    // likelihood kernel: hypotheses.  set s flips set_n[s] indicators; a set with no flip writes lncur.
    int nsets;
    int refresh_set;             // step kernel: the hypothesis set whose threads also bring te up to date (the lighter one)
    int prev_n;                  // step kernel: the chain kernels that may still be running when this launch starts sample
    int prev_person[4];          //   these prev_n (0-4) meioses (-1: the predecessor is something else, nothing may run
    int prev_parent[4];          //   ahead of it)
    int run_ahead;               // 2: two likelihood launches in flight (see slk_ms_step_kernel); else one
    // forest snapshots (step kernel, host record only): two buffers of slk_ms_snap_words(F) x M words
    uint32_t* snap;
    int snap_use, snap_build;    // buffer this launch starts its walks from / an extra set of CTAs fills (-1: none)
    int set_n[SLK_MS_MAXSETS];
    int set_person[SLK_MS_MAXSETS][2];
    int set_parent[SLK_MS_MAXSETS][2];
    // chain kernel: steps
    int nsteps;
    int step_person[2], step_parent[2];
    uint32_t step_slot[2];
    uint64_t seed, iteration;
    uint32_t chain;
    int32_t* dump_edges;         // optional [M][2N] (set 0)
    long long* trace;            // optional: clock64() stamps (tuning aid), see slk_debug_msampler_trace
    unsigned long long* timeline;// optional: %globaltimer stamps of CTA 0, 8 per launch (slk_debug_msampler_timeline)
    int tl_slot, tl_cta_off;     // tl_cta_off > 0: also (start, end) of every CTA of this launch at timeline[tl_cta_off + 2 * cta]
    double* out;                 // dg likelihood: [0] = sum ln(lik), [1] = recombination term
    const double* log_theta;
    const double* log_1mtheta;
    // step kernel: host-built record of this launch (slk_ms_pair_rec_words words; rec_n = 0: none, the kernel derives
    // it).  It travels in the kernel parameters: no copy to order against the launches, no buffer to keep alive.
    int rec_n;
    uint32_t rec[SLK_MS_REC_WORDS];
};

// Programmatic dependent launch (sm_90+): the step and chain kernels of a sweep alternate on one stream, each
// consuming what the previous one wrote.  Launched with cudaLaunchAttributeProgrammaticStreamSerialization, a
// kernel's CTAs become resident as soon as its predecessor has executed ms_launch_dependents(), stage their
// launch-invariant tables and then block in ms_wait_for_predecessor() until the predecessor grid has completed and
// its writes are visible: launch latency and prologue leave the dependent chain of a sweep.  Both are no-ops for
// a kernel launched without the attribute.
__device__ __forceinline__ unsigned long long ms_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define MS_TL(i) do { if(L.timeline && blockIdx.x == 0 && threadIdx.x == 0) L.timeline[8 * L.tl_slot + (i)] = ms_globaltimer(); } while(0)
__device__ __forceinline__ void ms_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void ms_wait_for_predecessor() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- per-thread arrays in shared memory, interleaved so that lane t always hits bank t ---------

struct MsLane {
    uint32_t base;               // shared-space byte address of the warp's slab + 4 * lane
    __device__ __forceinline__ uint32_t a8(uint32_t off, uint32_t k) const { return base + off + (((k & ~3u) << 5) | (k & 3u)); }
    __device__ __forceinline__ uint32_t a16(uint32_t off, uint32_t k) const { return base + off + (((k & ~1u) << 6) | ((k & 1u) << 1)); }
    __device__ __forceinline__ uint32_t a32(uint32_t off, uint32_t k) const { return base + off + (k << 7); }
};

// the "memory" clobbers keep every access in program order (the arrays alias each other freely)
__device__ __forceinline__ uint32_t ms_ld8(uint32_t a)  { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t ms_ld16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t ms_ld32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void ms_st8(uint32_t a, uint32_t v)  { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void ms_st16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void ms_st32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }

// Host and device agree on the carve-up through this function.  Per thread, in bytes:
//   keep     te (labels of typed people, 2*nt entries of E bytes) | obs (their genotypes, 2 bits each)
//   overlay  phase A: labels (2(N-F) entries of E bytes, founders are implicit) | graph row (N bytes)
//            phase C: grp (G node words, a second word each if wide)
// E = 1 if 2F <= 256 else 2 ("wide"); G = 2F: one disjoint-set node per founder allele (ms_walk).
struct MsLayout {
    int wide;
    int G;
    uint32_t te, obs, ov;        // region offsets for one thread (multiply by 32 lanes)
    uint32_t lab, row;           // phase A inside the overlay
    uint32_t grp, cnt, fa;       // phase C inside the overlay (cnt: second word per component, wide only)
    uint32_t per_thread;         // bytes per thread
    uint32_t cta_tables;         // bytes of CTA-shared tables (per-person records, typed list, x-male flags)
};

#if defined(__CUDACC__)
__host__ __device__
#endif
static inline MsLayout slk_ms_layout(int N, int F, int nt) {
    MsLayout L;
    L.wide = (2 * F > 255) ? 1 : 0;
    L.G = 2 * F;                     // one union-find word per founder allele (ms_walk)
    if(L.G < 1) L.G = 1;
    const uint32_t E = L.wide ? 2u : 1u;
#define r4(x) ((((uint32_t)(x)) + 3u) & ~3u)
#define r8(x) ((((uint32_t)(x)) + 7u) & ~7u)
    L.te = 0;
    L.obs = r4(2u * nt * E);
    L.ov = L.obs + 4u * ((nt + 15u) / 16u);
    L.lab = 0;
    L.row = L.lab + r4(2u * (N - F) * E);
    const uint32_t a_bytes = L.row + r4((uint32_t) N);
    L.grp = 0;
    L.cnt = L.grp + 4u * L.G;
    L.fa = L.cnt + (L.wide ? 4u * L.G : 0u);
    const uint32_t c_bytes = L.fa;
    L.per_thread = L.ov + r4(a_bytes > c_bytes ? a_bytes : c_bytes);
    L.cta_tables = r8((uint32_t)(16 * (N - F) + 4 * nt + nt + 8));
#undef r4
#undef r8
    return L;
}

// ---- (1) founder allele graph likelihood ---------------------------------------------------------
//
// The walk over the typed people (founder_allele_graph4.cc:34-424) shared by the two likelihood kernels, as a
// disjoint-set forest over the FOUNDER ALLELES with parity bits.
//
// A typed person is an edge between the founder alleles its two alleles descend from.  A heterozygous genotype says
// the two alleles differ, a homozygous one that both equal the observed allele; an assignment of A / B to the founder
// alleles has probability prod p(allele), and the likelihood is the sum over the assignments consistent with every
// edge.  It factorises over the connected components; a component is inconsistent (likelihood 0), or has one
// consistent assignment (some homozygous edge fixed it), or exactly two, complementary ones (only heterozygous edges).
// The reference keeps explicit component tables and relabels every founder allele on each merge
// (combine_components, founder_allele_graph4.cc:504-546); here every founder allele is a node of a disjoint-set
// forest: a word holds its parent, the parity to the parent ("my value is the opposite of my parent's") and, at a
// root, whether the component is fixed and to which root value, and n0 / n1 = how many of its alleles have the root's
// value / the opposite one.  A union is two stores; the counts play the role of the reference's two probability
// products: fixed to root value x the component weighs p(x)^n0 p(!x)^n1, unfixed M^n0 m^n1 + m^n0 M^n1, so the kernel
// returns ln L = A ln(major) + B ln(minor) + sum over unfixed components of ln(1 + (minor/major)^|n0-n1|): exact
// integer bookkeeping, one rounding step at the end (checked against the oracle's explicit products on every test
// pedigree).  An allele no typed person carries is a component of its own that sums to major + minor = 1.
//
// node word: bits 0..11 parent, bit 12 parity to parent, bits 13..14 state (0 unfixed, 1 / 2 fixed to root value
// 0 / 1), bit 15 touched by an edge, bits 16..23 n0, 24..31 n1 (wide: n0, n1 16 bits each in a second word).
// `src.next(k, g, mat, pat)` supplies the k-th typed person's observed genotype and founder-allele labels.
// Snapshots (step kernel, see there): `snap_in` = the forest after a set of typed people nobody touches for a while has
// been walked, saved by an earlier launch -- the walk starts from it instead of from singletons; `snap_out` = save the
// forest after this walk instead of finishing it.  Layout [word][M] (a lane is a locus: coalesced), words 0 .. nn - 1 the
// nodes (wide: nn more with the counts), then one word that is non-zero if the saved forest is inconsistent.
template<bool WIDE, class Src>
__device__ __forceinline__ double ms_walk(const MsLane& ln, uint32_t o_grp, uint32_t o_cnt, uint32_t /*o_fa*/, int F, int nt,
                                          const uint8_t* /*s_auto*/, Src& src, double lnM, double lnm, long long* tr,
                                          const uint32_t* snap_in = 0, uint32_t* snap_out = 0, size_t snap_stride = 0) {
    const int nn = 2 * F;
    bool dead = false;
    if(snap_in) {
        // (L2 loads: the buffer was last written by another launch, this SM's L1 may still hold what it held before)
        for(int i0 = 0; i0 < nn; i0 += 8) {
            uint32_t w[8], c[8];
#pragma unroll
            for(int j = 0; j < 8; ++j) if(i0 + j < nn) {
                w[j] = __ldcg(snap_in + (size_t)(i0 + j) * snap_stride);
                if(WIDE) c[j] = __ldcg(snap_in + (size_t)(nn + i0 + j) * snap_stride);
            }
#pragma unroll
            for(int j = 0; j < 8; ++j) if(i0 + j < nn) {
                ms_st32(ln.a32(o_grp, i0 + j), w[j]);
                if(WIDE) ms_st32(ln.a32(o_cnt, i0 + j), c[j]);
            }
        }
        dead = __ldcg(snap_in + (size_t)(WIDE ? 2 * nn : nn) * snap_stride) != 0u;
    }
    else {
        for(int i = 0; i < nn; ++i) {
            if(WIDE) { ms_st32(ln.a32(o_grp, i), (uint32_t) i); ms_st32(ln.a32(o_cnt, i), 1u); }
            else ms_st32(ln.a32(o_grp, i), (uint32_t) i | (1u << 16));
        }
    }
    struct Node { uint32_t meta, n0, n1; };
    auto node_ld = [&](uint32_t g) -> Node {
        Node r;
        const uint32_t w = ms_ld32(ln.a32(o_grp, g));
        r.meta = w & 0xffffu;
        if(WIDE) { const uint32_t c = ms_ld32(ln.a32(o_cnt, g)); r.n0 = c & 0xffffu; r.n1 = c >> 16; }
        else { r.n0 = (w >> 16) & 0xffu; r.n1 = w >> 24; }
        return r;
    };
    auto node_st = [&](uint32_t g, uint32_t meta, uint32_t n0, uint32_t n1) {
        if(WIDE) { ms_st32(ln.a32(o_grp, g), meta); ms_st32(ln.a32(o_cnt, g), n0 | (n1 << 16)); }
        else ms_st32(ln.a32(o_grp, g), meta | (n0 << 16) | (n1 << 24));
    };
    // root of node g (its entry e is passed in so that two first loads can be issued together): returns the root,
    // its entry, and the parity of g to the root
    auto find_from = [&](uint32_t g, Node e, Node& re, uint32_t& par) -> uint32_t {
        const uint32_t g0 = g;
        par = 0;
        int hops = 0;
        while((e.meta & 0xfffu) != g) {
            par ^= (e.meta >> 12) & 1u;
            g = e.meta & 0xfffu;
            e = node_ld(g);
            ++hops;
        }
        if(hops > 1) node_st(g0, g | (par << 12), 0u, 0u);           // path compression (a non-root's counts are dead)
        re = e;
        return g;
    };

    const int niter = src.iterations(nt);
    for(int it = 0; it < niter && !dead; ++it) {
        uint32_t g, mat, pat;
        bool single;                                    // x-linked male: the maternal allele only
        const int k = src.next(it, g, mat, pat, single);       // < 0: nobody to visit in this iteration
        if(k < 0 || g == SLK_UNTYPED) continue;
        const uint32_t gB = (g == SLK_HOMOZ_B) ? 1u : 0u;
        const bool het = g == SLK_HETERO;
        const bool self = mat == pat || single;
        if(self && het) { dead = true; break; }
        const Node eu = node_ld(mat), ev = node_ld(self ? mat : pat);      // both loads in flight together
        Node ru, rv; uint32_t pu, pv;
        const uint32_t u = find_from(mat, eu, ru, pu);
        uint32_t root = u, proot = pu;                   // the component's root after this edge, mat's parity to it
        Node rr = ru;
        if(!self) {
            const uint32_t v = find_from(pat, ev, rv, pv);
            const uint32_t pi = (het ? 1u : 0u) ^ pu ^ pv;        // required: value(root u) ^ value(root v)
            if(u == v) {
                if(pi) { dead = true; break; }
            }
            else {
                // union by size: the smaller tree hangs under the larger one's root
                const bool u_big = ru.n0 + ru.n1 >= rv.n0 + rv.n1;
                const uint32_t big = u_big ? u : v, small = u_big ? v : u;
                const Node rb = u_big ? ru : rv, rs = u_big ? rv : ru;
                uint32_t sb = (rb.meta >> 13) & 3u;
                const uint32_t ss = (rs.meta >> 13) & 3u;
                if(ss) {
                    const uint32_t xs = ((ss - 1u) ^ pi) + 1u;           // the small root's fixed value in the big root's frame
                    if(sb == 0u) sb = xs;
                    else if(sb != xs) { dead = true; break; }
                }
                rr.meta = big | (sb << 13) | 0x8000u;
                rr.n0 = rb.n0 + (pi ? rs.n1 : rs.n0);
                rr.n1 = rb.n1 + (pi ? rs.n0 : rs.n1);
                node_st(small, big | (pi << 12), 0u, 0u);
                root = big;
                proot = u_big ? pu : (pu ^ pi);
            }
        }
        uint32_t st = (rr.meta >> 13) & 3u;
        if(!het) {
            // homozygous: mat's allele is gB, i.e. the root's value is gB ^ parity(mat -> root)
            const uint32_t x = (gB ^ proot) + 1u;
            if(st == 0u) st = x;
            else if(st != x) { dead = true; break; }
        }
        node_st(root, root | (st << 13) | 0x8000u, rr.n0, rr.n1);
    }

    if(tr) tr[5] = clock64();
    if(snap_out) {
        for(int i = 0; i < nn; ++i) {
            snap_out[(size_t) i * snap_stride] = ms_ld32(ln.a32(o_grp, i));
            if(WIDE) snap_out[(size_t)(nn + i) * snap_stride] = ms_ld32(ln.a32(o_cnt, i));
        }
        snap_out[(size_t)(WIDE ? 2 * nn : nn) * snap_stride] = dead ? 1u : 0u;
        return 0.0;
    }
    double ret = -INFINITY;
    if(!dead) {
        uint32_t A = 0, B = 0;
        const bool minor_smaller = lnm <= lnM;
        const double rho = exp(-fabs(lnm - lnM));
        double prod = 1.0;
        for(int i0 = 0; i0 < nn; i0 += 4) {
            // four table words in flight at a time (the loads are ordered asm statements: one per iteration would
            // pay a shared-memory latency each)
            Node eq[4];
#pragma unroll
            for(int j = 0; j < 4; ++j) {
                if(i0 + j < nn) eq[j] = node_ld(i0 + j);
                else { eq[j].meta = 0xfffu; eq[j].n0 = eq[j].n1 = 0u; }
            }
#pragma unroll
            for(int j = 0; j < 4; ++j) {
                const Node e = eq[j];
                // roots that an edge touched only (an untouched allele sums to 1; a non-root is counted at its root)
                if((e.meta & 0xfffu) != (uint32_t)(i0 + j) || !(e.meta & 0x8000u)) continue;
                const uint32_t stt = (e.meta >> 13) & 3u;
                uint32_t a, b;
                if(stt == 0u) {
                    // two complementary assignments: M^n0 m^n1 + m^n0 M^n1; keep the larger term, more factors of the
                    // more frequent allele, and multiply up the (1 + rho^k) corrections
                    const uint32_t hi = e.n0 > e.n1 ? e.n0 : e.n1, lo = e.n0 > e.n1 ? e.n1 : e.n0;
                    uint32_t kk = hi - lo;
                    if(minor_smaller) { a = hi; b = lo; } else { a = lo; b = hi; }
                    double t = 1.0, sq = rho;
                    while(kk) { if(kk & 1u) t *= sq; sq *= sq; kk >>= 1; }
                    prod *= 1.0 + t;
                }
                else if(stt == 1u) { a = e.n0; b = e.n1; }       // root value 0 (allele A, the major allele's exponent first)
                else               { a = e.n1; b = e.n0; }
                A += a; B += b;
            }
        }
        ret = ((double) A * lnM + (double) B * lnm) + log(prod);
        if(!(ret > -1e290)) ret = -INFINITY;               // a zero allele frequency entered the product
    }
    return ret;
}

template<bool WIDE>
__global__ void __launch_bounds__(32)
slk_ms_likelihood_kernel(const SlkMsLaunch L) {
    extern __shared__ __align__(16) unsigned char ms_smem[];
    const int N = L.N, F = L.F, M = L.M, nt = L.ms.n_typed;
    const MsLayout lay = slk_ms_layout(N, F, nt);
    const uint32_t t = threadIdx.x;
    long long* tr = (L.trace && t == 0 && (blockIdx.x & 63) == 0 && (blockIdx.x >> 6) < 12) ? L.trace + 8 * (blockIdx.x >> 6) : 0;
    if(tr) tr[0] = clock64();

    // the phase A and phase C tables overlay each other ACROSS the lanes of the warp, so the phases are
    // separated by __syncwarp()
    MsLane ln;
    ln.base = (uint32_t) __cvta_generic_to_shared(ms_smem + lay.cta_tables) + (t << 2);
    const uint32_t o_te = lay.te * 32u, o_obs = lay.obs * 32u, o_ov = lay.ov * 32u;
    const uint32_t o_lab = o_ov + lay.lab * 32u, o_row = o_ov + lay.row * 32u;
    const uint32_t o_grp = o_ov + lay.grp * 32u, o_cnt = o_ov + lay.cnt * 32u, o_fa = o_ov + lay.fa * 32u;

    // byte offset (inside the lane-interleaved label array) of the label PAIR of non-founder p: the two
    // labels are adjacent (narrow: two bytes of one 16-bit half-word; wide: one 32-bit word)
#define PAIR_OFF(p) (WIDE ? ln_a16_off(2u * (uint32_t)((p) - F)) : ln_a8_off(2u * (uint32_t)((p) - F)))
    auto ln_a8_off = [](uint32_t k) -> uint32_t { return ((k & ~3u) << 5) | (k & 3u); };
    auto ln_a16_off = [](uint32_t k) -> uint32_t { return ((k & ~1u) << 6) | ((k & 1u) << 1); };

    // CTA-shared tables, built once per warp: for the k-th non-founder in topological order
    //   x = person << 16 | offset of its graph-row byte, y = offset of its label pair,
    //   z / w = mother's / father's label-pair offset, or 0x80000000 | 2 * parent for a founder parent;
    // for the k-th typed person its label-pair offset or 0x80000000 | 2 * person
    uint4* s_rec = (uint4*) ms_smem;                       // [N-F]
    uint32_t* s_trec = (uint32_t*)(s_rec + (N - F));       // [nt]
    uint8_t* s_auto = (uint8_t*)(s_trec + nt);             // [nt] x-linked male: maternal allele only
    for(int k = t; k < N - F; k += 32) {
        const int i = L.ms.seq[k], mo = L.mother[i], fa = L.father[i];
        uint4 r;
        r.x = ((uint32_t) i << 16) | ln_a8_off((uint32_t) i);
        r.y = PAIR_OFF(i);
        r.z = mo < F ? (0x80000000u | (uint32_t)(2 * mo)) : PAIR_OFF(mo);
        r.w = fa < F ? (0x80000000u | (uint32_t)(2 * fa)) : PAIR_OFF(fa);
        s_rec[k] = r;
    }
    for(int k = t; k < nt; k += 32) {
        const int p = L.ms.typed[k];
        s_trec[k] = p < F ? (0x80000000u | (uint32_t)(2 * p)) : PAIR_OFF(p);
        s_auto[k] = (L.sex_linked && L.male[p]) ? 1 : 0;
    }
    __syncwarp();

    const int nblk = (M + 31) >> 5;
    const int set = blockIdx.x / nblk;
    const int l = (blockIdx.x - set * nblk) * 32 + (int) t;
    const bool live = l < M;
    const int nflip = L.set_n[set];
    const int p0 = nflip > 0 ? L.set_person[set][0] : -1, a0 = L.set_parent[set][0];
    const int p1 = nflip > 1 ? L.set_person[set][1] : -1, a1 = L.set_parent[set][1];

    if(tr) tr[1] = clock64();
    // ---- phase A: labels -----------------------------------------------------------------------
    if(live) {
        const uint8_t* row = L.dgp + (size_t) l * N;
        if(((N | (int)(size_t) L.dgp) & 3) == 0) {
            const uint32_t* row4 = (const uint32_t*) row;
            for(int i = F >> 2; i < (N >> 2); ++i) ms_st32(ln.a32(o_row, i), __ldg(row4 + i));
        }
        else {
            for(int i = F; i < N; ++i) ms_st8(ln.a8(o_row, i), row[i]);
        }
        // observed genotypes of the typed people ([nt][M] in global memory: coalesced over the lanes),
        // sixteen independent loads in flight, packed two bits each
        {
            const uint8_t* obs = L.ms.obsT + l;
            for(int w = 0; w * 16 < nt; ++w) {
                uint32_t acc = 0;
#pragma unroll
                for(int j = 0; j < 16; ++j) {
                    const int k = w * 16 + j;
                    if(k < nt) acc |= (uint32_t) __ldg(obs + (size_t) k * M) << (2 * j);
                }
                ms_st32(ln.a32(o_obs, w), acc);
            }
        }
        if(tr) tr[2] = clock64();
        const uint32_t lab_base = ln.base + o_lab, row_base = ln.base + o_row;
        for(int k = 0; k < N - F; ++k) {
            const uint4 r = s_rec[k];
            const int i = (int)(r.x >> 16);
            uint32_t b = ms_ld8(row_base + (r.x & 0xffffu));
            if(i == p0) {
                if(set < 2) L.bits[(size_t) set * M + l] = (uint8_t)((b >> a0) & 1u);
                b ^= (1u << a0);
            }
            if(i == p1) b ^= (1u << a1);
            const uint32_t bm = b & 1u, bf = (b >> 1) & 1u;
            uint32_t m, f;
            if(WIDE) {
                m = (r.z & 0x80000000u) ? (r.z & 0xffffu) + bm : ms_ld16(lab_base + r.z + 2u * bm);
                f = (r.w & 0x80000000u) ? (r.w & 0xffffu) + bf : ms_ld16(lab_base + r.w + 2u * bf);
                ms_st32(lab_base + r.y, m | (f << 16));
            }
            else {
                m = (r.z & 0x80000000u) ? (r.z & 0xffffu) + bm : ms_ld8(lab_base + r.z + bm);
                f = (r.w & 0x80000000u) ? (r.w & 0xffffu) + bf : ms_ld8(lab_base + r.w + bf);
                ms_st16(lab_base + r.y, m | (f << 8));
            }
        }
        if(L.dump_edges && set == 0) {
            int32_t* e = L.dump_edges + (size_t) l * 2 * N;
            for(int i = 0; i < 2 * F; ++i) e[i] = i;
            for(int i = 2 * F; i < 2 * N; ++i)
                e[i] = (int32_t)(WIDE ? ms_ld16(ln.a16(o_lab, i - 2 * F)) : ms_ld8(ln.a8(o_lab, i - 2 * F)));
        }
        if(tr) tr[3] = clock64();
        // ---- phase B: keep the typed people's label pairs ------------------------------------------
        for(int k = 0; k < nt; ++k) {
            const uint32_t c = s_trec[k];
            if(WIDE) {
                const uint32_t v = (c & 0x80000000u) ? ((c & 0xffffu) | (((c & 0xffffu) + 1u) << 16)) : ms_ld32(lab_base + c);
                ms_st32(ln.a16(o_te, 2 * k), v);
            }
            else {
                const uint32_t v = (c & 0x80000000u) ? ((c & 0xffu) | (((c & 0xffu) + 1u) << 8)) : ms_ld16(lab_base + c);
                ms_st16(ln.a8(o_te, 2 * k), v);
            }
        }
    }
    __syncwarp();                                      // every lane is done with its labels
    if(!live) return;
    double ret;
    {
        struct SmemSource {
            const MsLane& ln; uint32_t o_te, o_obs; uint32_t gw; const uint8_t* s_auto;
            __device__ __forceinline__ int iterations(int nt) const { return nt; }
            __device__ __forceinline__ int next(int k, uint32_t& g, uint32_t& mat, uint32_t& pat, bool& single) {
                single = s_auto[k] != 0;
                if((k & 15) == 0) gw = ms_ld32(ln.a32(o_obs, k >> 4));
                g = (gw >> (2 * (k & 15))) & 3u;
                if(WIDE) { const uint32_t v = ms_ld32(ln.a16(o_te, 2 * k)); mat = v & 0xffffu; pat = v >> 16; }
                else { const uint32_t v = ms_ld16(ln.a8(o_te, 2 * k)); mat = v & 0xffu; pat = v >> 8; }
                return k;
            }
        } src = { ln, o_te, o_obs, 0u, s_auto };
        if(tr) tr[4] = clock64();
        ret = ms_walk<WIDE>(ln, o_grp, o_cnt, o_fa, F, nt, s_auto, src, L.ms.lnmajor[l], L.ms.lnminor[l], tr);
    }
    if(nflip == 0 && L.te) {
        // reset: publish the typed people's label pairs of the current graph (the incremental step kernel
        // reads them) and clear the locus's stale mask
        for(int k = 0; k < nt; ++k) {
            if(WIDE) ((uint32_t*) L.te)[(size_t) k * M + l] = ms_ld32(ln.a16(o_te, 2 * k));
            else     ((uint16_t*) L.te)[(size_t) k * M + l] = (uint16_t) ms_ld16(ln.a8(o_te, 2 * k));
        }
        for(int w = 0; w < L.ms.W; ++w) L.stale[(size_t) l * L.ms.W + w] = 0u;
    }
    if(nflip == 0) L.lncur[l] = ret;
    else L.lnl[(size_t) set * M + l] = ret;
    if(tr) tr[6] = clock64();
#undef PAIR_OFF
}

// ---- (1b) the same likelihood, incrementally: only the labels a hypothesis can change are recomputed ------
//
// Between two steps of a sweep the graph changes by one indicator per locus at most, and flipping the
// indicator of person P can only change the labels of P's own slot and of slots below P.  The chain keeps
// the typed people's label pairs of the CURRENT graph in global memory (te[nt][M], written by the full
// kernel at the start of a sweep).  A hypothesis thread reads them (one person ahead of their use, coalesced
// over the lanes) and re-derives a label only where the hypothesis's slot mask says so, by walking the
// lineage up the graph row to a founder (a handful of byte loads).  The slot mask is
//     below(P_0) | below(P_1) | own slots | stale[l],
// where stale[l] collects the slots invalidated by indicators the chain kernel has flipped at this locus since
// te was last brought up to date; the set-0 thread of the locus writes those entries back at the end of its
// work and the chain kernel clears the mask before it records new flips.  No per-thread label array, no
// graph-row or genotype staging: 16 KB of shared memory per warp instead of 31.
struct MsStepLayout {
    int wide, G;
    uint32_t grp, cnt, fa, hmask;    // per-thread byte offsets (multiply by 32 lanes)
    uint32_t per_thread, cta_tables;
};

#if defined(__CUDACC__)
__host__ __device__
#endif
static inline MsStepLayout slk_ms_step_layout(int N, int F, int nt, int W) {
    MsStepLayout L;
    L.wide = (2 * F > 255) ? 1 : 0;
    L.G = 2 * F;
    if(L.G < 1) L.G = 1;
    L.grp = 0;
    L.cnt = L.grp + 4u * L.G;
    L.fa = L.cnt + (L.wide ? 4u * L.G : 0u);
    L.hmask = L.fa;
    L.per_thread = L.hmask + 4u * W;
    L.cta_tables = ((uint32_t)(4 * N + 2 * nt + nt + 4 + 4 * W + 2 * nt + 8 + 8) + 7u) & ~7u;
    return L;
}

// Host-built record of one launch of a sweep (it travels in the kernel parameters): word 0 = number of visited typed
// people the running chain kernels cannot affect (the phase boundary), word 1 = people visited in all (everybody, or
// everybody a snapshot does not already hold), word 2 = people the snapshot set visits; then the slot masks of the three
// hypotheses (3 x W words, without the locus's stale slots), then the visiting order, two 16-bit entries per word
// (index | single-allele flag << 15), then the snapshot set's order.  With a record the per-warp prologue is one batch of
// independent loads.
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline int slk_ms_pair_rec_words(int nt, int W) { return 3 + 3 * W + 2 * ((nt + 1) / 2); }

// words per locus of a forest snapshot (ms_walk): the nodes (two words each if wide) and the inconsistency flag
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline int slk_ms_snap_words(int F) { return ((2 * F > 255) ? 4 * F : 2 * F) + 1; }

template<bool WIDE>
__global__ void __launch_bounds__(32)
slk_ms_step_kernel(const SlkMsLaunch L) {
    extern __shared__ __align__(16) unsigned char ms_smem[];
    const int N = L.N, F = L.F, M = L.M, nt = L.ms.n_typed, W = L.ms.W;
    const MsStepLayout lay = slk_ms_step_layout(N, F, nt, W);
    const uint32_t t = threadIdx.x;
    long long* tr = (L.trace && t == 0 && (blockIdx.x & 63) == 0 && (blockIdx.x >> 6) < 12) ? L.trace + 8 * (blockIdx.x >> 6) : 0;
    if(tr) tr[0] = clock64();
    MS_TL(0);
    if(L.timeline && L.tl_cta_off > 0 && t == 0) L.timeline[L.tl_cta_off + 2 * blockIdx.x] = ms_globaltimer() & 0xffffffffffffull;

    // Run-ahead 1: the successor (this pair's chain kernel) may become resident at once.  Run-ahead 2: only after this
    // launch's own predecessor has completed (in sync()), which is what bounds the pipeline to two likelihood launches.
    const bool deep = L.run_ahead == 2;
    if(!deep) ms_launch_dependents();
    int16_t* s_mo = (int16_t*) ms_smem;                    // [N]
    int16_t* s_fa = s_mo + N;                              // [N]
    uint16_t* s_typed = (uint16_t*)(s_fa + N);             // [nt]
    uint8_t* s_auto = (uint8_t*)(s_typed + nt);            // [nt]
    for(int i = t; i < N; i += 32) { s_mo[i] = L.mother[i]; s_fa[i] = L.father[i]; }
    uint32_t* s_prev = (uint32_t*)(ms_smem + ((4u * N + 3u * nt + 3u) & ~3u));      // [W]
    uint16_t* s_ord = (uint16_t*)(s_prev + W);                                      // [nt]
    int* s_n0 = (int*)(ms_smem + ((4u * N + 3u * nt + 3u) & ~3u) + 4u * W + ((2u * nt + 3u) & ~3u));
    const bool rec = L.rec_n > 0;
    const int nblk = (M + 31) >> 5;
    // Snapshot set (an extra set of CTAs, host record only): walks the typed people that no meiosis of the next few
    // launches can affect and saves the forest; those launches start their walks from it.  It takes the FIRST block
    // indices: the chain kernel that follows waits for every CTA of this launch, and the snapshot set's walk is the
    // longest; the hypothesis sets have slack before their wait (7.94 -> 7.90 ms per sweep against taking the last).
    const int bset = blockIdx.x / nblk;
    const bool has_build = L.snap_build >= 0;
    const bool build = has_build && bset == 0;
    const int set = has_build ? (build ? L.nsets : bset - 1) : bset;
    const int n_visit = rec ? (int) L.rec[build ? 2 : 1] : nt;
    if(rec) {
        // the order and the phase boundary come from the host
        const uint32_t* ordw = L.rec + 3 + 3 * W + (build ? (nt + 1) / 2 : 0);
        for(int k = t; k < nt; k += 32) s_typed[k] = L.ms.typed[k];
        for(int k = t; k < (nt + 1) / 2; k += 32) ((uint32_t*) s_ord)[k] = ordw[k];
        if(t == 0) *s_n0 = build ? n_visit : (int) L.rec[0];
    }
    else {
    for(int k = t; k < nt; k += 32) {
        const int p = L.ms.typed[k];
        s_typed[k] = (uint16_t) p;
        s_auto[k] = (L.sex_linked && L.male[p]) ? 1 : 0;
    }
    // slots the predecessor kernel's flips can change (its two meioses: own slot and everything below; all of them
    // if the predecessor is not a chain kernel), and the visiting order of the walk: first the typed people who
    // have none of those slots (ascending), then the others (ascending)
    for(int w = t; w < W; w += 32) {
        uint32_t m = L.prev_n < 0 ? 0xffffffffu : 0u;
        for(int j = 0; j < L.prev_n; ++j) {
            m |= L.ms.desc_mask[(size_t)(L.prev_person[j] - F) * W + w];
            const int kk = L.ms.typed_index[L.prev_person[j]];
            if(kk >= 0 && ((2 * kk + L.prev_parent[j]) >> 5) == w) m |= 1u << ((2 * kk + L.prev_parent[j]) & 31);
        }
        s_prev[w] = m;
    }
    __syncwarp();
    {
        int n = 0;
        for(int pass = 0; pass < 2; ++pass) {
            for(int k0 = 0; k0 < nt; k0 += 32) {
                const int k = k0 + (int) t;
                const bool mine = k < nt && ((((s_prev[k >> 4] >> (2 * (k & 15))) & 3u) != 0) == (pass == 1));
                const uint32_t b = __ballot_sync(0xffffffffu, mine);
                if(mine) s_ord[n + __popc(b & ((1u << t) - 1u))] = (uint16_t)(k | (s_auto[k] ? 0x8000 : 0));     // N < 2048
                n += __popc(b);
            }
            if(pass == 0 && t == 0) *s_n0 = n;
        }
    }
    }
    __syncwarp();

    const int l = (blockIdx.x - bset * nblk) * 32 + (int) t;
    if(l >= M) return;
    const int hset = build ? 0 : set;
    const int nflip = build ? 0 : L.set_n[hset];
    const int p0 = nflip > 0 ? L.set_person[hset][0] : -1, a0 = L.set_parent[hset][0];
    const int p1 = nflip > 1 ? L.set_person[hset][1] : -1, a1 = L.set_parent[hset][1];

    MsLane ln;
    ln.base = (uint32_t) __cvta_generic_to_shared(ms_smem + lay.cta_tables) + (t << 2);
    const uint32_t o_grp = lay.grp * 32u, o_cnt = lay.cnt * 32u, o_fa = lay.fa * 32u, o_hm = lay.hmask * 32u;
    const uint8_t* row = L.dgp + (size_t) l * N;
    // the hypothesis's slot mask, without the locus's stale slots: those are the predecessor kernel's to write
    if(rec) {
        for(int w = 0; w < W; ++w) ms_st32(ln.a32(o_hm, w), build ? 0u : L.rec[3 + set * W + w]);
    }
    else {
    for(int w = 0; w < W; ++w) {
        uint32_t h = 0;
        if(p0 >= 0) h |= L.ms.desc_mask[(size_t)(p0 - F) * W + w];
        if(p1 >= 0) h |= L.ms.desc_mask[(size_t)(p1 - F) * W + w];
        ms_st32(ln.a32(o_hm, w), h);
    }
    if(p0 >= 0) {
        const int k0 = L.ms.typed_index[p0];
        if(k0 >= 0) { const uint32_t q = 2 * k0 + a0; ms_st32(ln.a32(o_hm, q >> 5), ms_ld32(ln.a32(o_hm, q >> 5)) | (1u << (q & 31))); }
    }
    if(p1 >= 0) {
        const int k1 = L.ms.typed_index[p1];
        if(k1 >= 0) { const uint32_t q = 2 * k1 + a1; ms_st32(ln.a32(o_hm, q >> 5), ms_ld32(ln.a32(o_hm, q >> 5)) | (1u << (q & 31))); }
    }
    }

    // The walk in two phases.  The predecessor on the stream (the chain kernel of the previous pair of steps) flips
    // indicators of that pair's two people only.  A typed person none of whose slots is one of theirs or below
    // them (s_prev) has the label pair and the lineages it had before that kernel ran, so phase 0 visits those
    // people -- 93 % of them on the 200-member pedigree -- while the predecessor is still sampling.  sync() then
    // blocks until the predecessor has completed, folds the locus's stale slots into the mask and records the
    // indicator's current value, and phase 1 visits the rest (s_ord is the whole order, *s_n0 the phase boundary).  Nothing is written to global memory before sync()
    // (the predecessor is still reading ln L and `bits` of ITS steps); te is only ever written by likelihood
    // kernels, so its entries are streamed ahead of their use in both phases.
    struct StepSource {
        const MsLane& ln; uint32_t o_hm; const uint8_t* row; const int16_t* s_mo; const int16_t* s_fa; const uint16_t* s_typed;
        const uint16_t* s_ord; int n0;
        const void* te; const uint8_t* obs; const uint32_t* stale; uint8_t* bits_out; int M, l, F, nt, W, p0, a0, p1, a1;
        bool synced; unsigned long long* tl; bool deep;
        uint32_t hw_next;
        uint32_t qk[4], qp[4], qg[4];    // the next four people of the order: index, label pair, genotype (loads in flight)
        __device__ __forceinline__ int iterations(int n) const { return n; }
        __device__ __forceinline__ void sync() {
            if(tl) tl[2] = ms_globaltimer();
            ms_wait_for_predecessor();
            if(deep) ms_launch_dependents();
            if(tl) tl[3] = ms_globaltimer();
            synced = true;
            for(int w = 0; w < W; ++w) {
                const uint32_t st = stale[(size_t) l * W + w];
                if(st) ms_st32(ln.a32(o_hm, w), ms_ld32(ln.a32(o_hm, w)) | st);
            }
            if(bits_out) *bits_out = (uint8_t)((row[p0] >> a0) & 1u);
        }
        // founder allele reached from slot `sl` of `person` by following the indicators of the graph row
        __device__ __forceinline__ uint32_t trace(int person, uint32_t sl, bool flips) const {
            int cur = person;
            while(cur >= F) {
                uint32_t b = row[cur];           // L1-resident after the first touch (staging the row in shared memory
                                                 // costs more in the prologue than these few reads save)
                if(flips) { if(cur == p0) b ^= 1u << a0; if(cur == p1) b ^= 1u << a1; }
                const uint32_t nb = (b >> sl) & 1u;
                cur = sl ? s_fa[cur] : s_mo[cur];
                sl = nb;
            }
            return (uint32_t)(2 * cur) + sl;
        }
        __device__ __forceinline__ void fetch(int it, int slot) {
            if(it < nt) {
                const uint32_t ko = s_ord[it], k = ko & 0x7fffu;
                qk[slot] = ko;
                qp[slot] = WIDE ? ((const uint32_t*) te)[(size_t) k * M + l] : (uint32_t)((const uint16_t*) te)[(size_t) k * M + l];
                qg[slot] = obs[(size_t) k * M];
            }
        }
        __device__ __forceinline__ int next(int it, uint32_t& g, uint32_t& mat, uint32_t& pat, bool& single) {
            if(it == n0) {                               // uniform: every lane of the launch is at the same iteration
                sync();
                // run-ahead 2: the previous likelihood launch was still refreshing the label pairs of the people
                // visited from here on when the queue read them ahead: read them again (after the wait, which makes
                // the completed grids' writes visible to this one)
                if(deep) { for(int j = 0; j < 4; ++j) fetch(it + j, j); }
                hw_next = ms_ld32(ln.a32(o_hm, (qk[0] & 0x7fffu) >> 4));
            }
            // Four-deep queues in registers (shifted, so every index is a compile-time constant; a ring indexed by
            // it & 3 through a switch measured 35 % slower: four copies of the body).  The person's index and mask
            // word are fetched one iteration ahead too: nothing of the order's indirection is on the walk's
            // dependent chain.
            const uint32_t k = qk[0] & 0x7fffu, pair = qp[0], hw = hw_next;
            single = (qk[0] & 0x8000u) != 0;
            g = qg[0];
            qk[0] = qk[1]; qp[0] = qp[1]; qg[0] = qg[1]; qk[1] = qk[2]; qp[1] = qp[2]; qg[1] = qg[2];
            qk[2] = qk[3]; qp[2] = qp[3]; qg[2] = qg[3];
            fetch(it + 4, 3);
            hw_next = ms_ld32(ln.a32(o_hm, (qk[0] & 0x7fffu) >> 4));
            const uint32_t m2 = (hw >> (2 * (k & 15))) & 3u;
            mat = WIDE ? (pair & 0xffffu) : (pair & 0xffu);
            pat = WIDE ? (pair >> 16) : (pair >> 8);
            if(m2 && g != SLK_UNTYPED) {
                const int person = s_typed[k];
                if(m2 & 1u) mat = trace(person, 0u, true);
                if(m2 & 2u) pat = trace(person, 1u, true);
            }
            return (int) k;
        }
    } src = { ln, o_hm, row, s_mo, s_fa, s_typed, s_ord, *s_n0, L.te, L.ms.obsT + l, L.stale,
              (p0 >= 0 && set < 2) ? L.bits + (size_t) set * M + l : (uint8_t*) 0,
              M, l, F, n_visit, W, p0, a0, p1, a1, false,
              (L.timeline && blockIdx.x == 0 && t == 0) ? L.timeline + 8 * L.tl_slot : (unsigned long long*) 0, deep,
              0u, {0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u} };
    const double lnM = L.ms.lnmajor[l], lnm = L.ms.lnminor[l];
    // predecessor unknown (the reset kernel, which writes te, or anything else): nothing is read ahead of it
    if(L.prev_n < 0) ms_wait_for_predecessor();
    for(int j = 0; j < 4; ++j) src.fetch(j, j);
    src.hw_next = ms_ld32(ln.a32(o_hm, (src.qk[0] & 0x7fffu) >> 4));
    if(tr) tr[4] = clock64();
    MS_TL(1);
    const size_t snap_size = (size_t) slk_ms_snap_words(F) * (size_t) M;
    const uint32_t* snap_in = (!build && L.snap_use >= 0) ? L.snap + (size_t) L.snap_use * snap_size + l : (const uint32_t*) 0;
    uint32_t* snap_out = build ? L.snap + (size_t) L.snap_build * snap_size + l : (uint32_t*) 0;
    const double ret = ms_walk<WIDE>(ln, o_grp, o_cnt, o_fa, F, n_visit, s_auto, src, lnM, lnm, tr, snap_in, snap_out, (size_t) M);
    // the snapshot set is done: it never waits for the predecessor (the people it visits are out of every running
    // kernel's reach) and its exit counts as its release of the dependent launch
    if(build) return;
    if(!src.synced) src.sync();                            // every lane left the walk (impossible graph) in phase 0
    L.lnl[(size_t) set * M + l] = ret;

    // bring the locus's out-of-date entries up to date (current graph: no hypothetical flips); the threads of the
    // other sets never use those entries (their masks contain the stale slots), so any set may do it: the host picks
    // the one whose hypothesis has fewer slots to re-derive
    if(set == L.refresh_set) {
        for(int w = 0; w < W; ++w) {
            uint32_t st = L.stale[(size_t) l * W + w];
            while(st) {
                const uint32_t q = 32u * w + (uint32_t) __ffs(st) - 1u;
                st &= st - 1u;
                const uint32_t k = q >> 1, sl = q & 1u;
                const uint32_t lab = src.trace(s_typed[k], sl, false);
                if(WIDE) ((uint16_t*) L.te)[2 * ((size_t) k * M + l) + sl] = (uint16_t) lab;
                else     ((uint8_t*) L.te)[2 * ((size_t) k * M + l) + sl] = (uint8_t) lab;
            }
        }
    }
    if(tr) tr[6] = clock64();
    MS_TL(4);
    __syncwarp();
    if(L.timeline && L.tl_cta_off > 0 && t == 0) {
        uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        L.timeline[L.tl_cta_off + 2 * blockIdx.x + 1] = ((ms_globaltimer() & 0xffffffffffffull) << 10) | (smid & 1023u);
    }
}

// ---- (2) + (3): forward pass and backward sampling along the chromosome ------------------------------

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
    return (u * (w0 + w1) < w0) ? 0 : 1;           // u < w0 / (w0 + w1) without the division
}

__device__ __forceinline__ Mat2 mat2_shfl_up(const Mat2& v, int d) {
    Mat2 r;
    r.a = __shfl_up_sync(0xffffffffu, v.a, d); r.b = __shfl_up_sync(0xffffffffu, v.b, d);
    r.c = __shfl_up_sync(0xffffffffu, v.c, d); r.d = __shfl_up_sync(0xffffffffu, v.d, d);
    return r;
}

// One thread-block CLUSTER of SLK_MS_CLUSTER CTAs (the chromosome is cut into contiguous chunks, one
// per thread of the cluster); the two scans cross CTAs through distributed shared memory.  Dynamic
// shared memory of a CTA, per locus of its share: raw pair (16 B), ln cur (8 B), ln flipped (8 B), theta of
// the interval to its left (8 B), the step's uniform draw (8 B), map byte, changed-in-step-0 byte.
#define SLK_MS_CLUSTER 8
#define SLK_MS_CHAIN_THREADS 640
#define SLK_MS_CHAIN_BYTES_PER_LOCUS 58

__global__ void __cluster_dims__(SLK_MS_CLUSTER, 1, 1) __launch_bounds__(SLK_MS_CHAIN_THREADS)
slk_ms_chain_kernel(const SlkMsLaunch L) {
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) unsigned char ch_smem[];
    __shared__ Mat2 s_wmat[32];
    __shared__ Mat2 s_cmat[SLK_MS_CLUSTER];        // the product of each CTA of the cluster
    __shared__ uint8_t s_wmap[32];
    __shared__ uint8_t s_cmap[SLK_MS_CLUSTER];
    __shared__ double2 s_r0;                       // raw pair of locus 0 (published by CTA 0)
    __shared__ Mat2 s_cpre;                        // product of the CTAs before this one
    __shared__ uint32_t s_cpost;                   // composition of the CTAs after this one
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int) cluster.block_rank();
    const int M = L.M, N = L.N, T = blockDim.x, t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5, nwarps = T >> 5;
    const int C = (M + SLK_MS_CLUSTER * T - 1) / (SLK_MS_CLUSTER * T);      // loci per thread
    const int cta_lo = min(M, rank * T * C), cta_hi = min(M, cta_lo + T * C);
    const int lo = min(M, cta_lo + t * C), hi = min(M, lo + C);
    const int cap = T * C;
    double2* s_raw = (double2*) ch_smem;                                    // [cap]
    double* s_lncur = (double*)(s_raw + cap);                               // [cap]
    double* s_lnflip = s_lncur + cap;                                       // [cap]
    double* s_theta = s_lnflip + cap;                                       // [cap] theta[i - 1]
    double* s_u = s_theta + cap;                                            // [2][cap]
    uint8_t* s_map = (uint8_t*)(s_u + 2 * cap);                             // [cap]
    uint8_t* s_changed = s_map + cap;                                       // [cap]

    auto apply = [](uint32_t f, uint32_t x) -> uint32_t { return (f >> x) & 1u; };
    auto compose = [&](uint32_t f, uint32_t g) -> uint32_t { return apply(f, apply(g, 0)) | (apply(f, apply(g, 1)) << 1); };

    long long* tr = (L.trace && t == 0 && (rank == 0 || rank == SLK_MS_CLUSTER - 1)) ? L.trace + (rank ? 24 : 0) : 0;
    if(tr) tr[0] = clock64();
    const int W = L.ms.W;
    MS_TL(0);
    // Run-ahead 2: the next pair's likelihood launch may start now, next to the launch this kernel waits for: it
    // visits nobody this kernel's or the previous chain kernel's meioses can affect before this kernel has completed.
    if(L.run_ahead == 2) ms_launch_dependents();
    // everything that does not depend on the likelihoods is done while the likelihood kernel is still running
    for(int i = cta_lo + t; i < cta_hi; i += T) {
        s_theta[i - cta_lo] = i > 0 ? L.theta[i - 1] : 0.5;
        for(int step = 0; step < L.nsteps; ++step)
            s_u[step * cap + i - cta_lo] = slk_uniform(L.seed, L.chain, L.iteration, (uint32_t) i, L.step_slot[step]);
    }
    MS_TL(1);
    ms_wait_for_predecessor();                             // ln L of the hypotheses, the stale masks' last readers
    MS_TL(2);
    // run-ahead 1, only now: the next step kernel reads te while it waits for this kernel, and the step kernel that has
    // just completed was still refreshing te (run-ahead 2 keeps the refreshed entries out of that launch's first phase)
    if(L.run_ahead != 2) ms_launch_dependents();
    for(int i = cta_lo + t; i < cta_hi; i += T) {
        s_lncur[i - cta_lo] = L.lncur[i]; s_changed[i - cta_lo] = 0;
        for(int w = 0; w < W; ++w) L.stale[(size_t) i * W + w] = 0u;       // the likelihood launch brought te up to date
    }
    const double theta_hi = cta_hi < M && cta_hi > 0 ? L.theta[cta_hi - 1] : 0.5;   // interval right of the CTA's last locus

    for(int step = 0; step < L.nsteps; ++step) {
        const uint8_t* bits = L.bits + (size_t) step * M;
        // raw_matrix of the step (meiosis_sampler.cc:117-131) up to a common factor per locus, which the
        // normalised recurrence does not see: [current value] = 1, [other] = L(flipped) / L(current).
        // In step 1 "flipped" is conditional on what step 0 sampled at this locus.
        for(int i = cta_lo + t; i < cta_hi; i += T) {
            const int k = i - cta_lo;
            const double lf = step == 0 ? L.lnl[i] : (s_changed[k] ? L.lnl[2 * (size_t) M + i] : L.lnl[(size_t) M + i]);
            const double lc = s_lncur[k];
            const double r = exp(lf - lc);                 // -inf -> 0
            s_lnflip[k] = lf;
            s_raw[k] = (bits[i] & 1u) ? make_double2(r, 1.0) : make_double2(1.0, r);
            if(!(lc > -DBL_MAX) && atomicCAS(&L.err[0], 0, SLK_ERR_ILLEGAL_GRAPH) == 0) L.err[1] = i;
        }
        __syncthreads();
        if(tr) tr[1 + 8 * step] = clock64();

        // ---- forward: product of the chunk's matrices A_i = diag(raw_i) * [[1-th, th], [th, 1-th]] ----
        Mat2 acc; acc.a = 1.0; acc.b = 0.0; acc.c = 0.0; acc.d = 1.0;
        for(int i = lo; i < hi; ++i) {
            if(i == 0) continue;                       // locus 0 enters as the start vector
            const double2 r = s_raw[i - cta_lo];
            const double th = s_theta[i - cta_lo], ith = 1.0 - th;
            Mat2 A; A.a = r.x * ith; A.b = r.x * th; A.c = r.y * th; A.d = r.y * ith;
            acc = mat2_mul_rescale(A, acc);
        }
        if(tr) tr[2 + 8 * step] = clock64();
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
            if(rank == 0) *cluster.map_shared_rank(&s_r0, t) = s_raw[0];       // locus 0's pair, to every CTA
        }
        if(tr) tr[3 + 8 * step] = clock64();
        cluster.sync();
        if(tr) tr[4 + 8 * step] = clock64();
        if(t == 0) {
            // product of the CTAs before this one, once per CTA
            Mat2 q; q.a = 1.0; q.b = 0.0; q.c = 0.0; q.d = 1.0;
            for(int c = 0; c < rank; ++c) q = mat2_mul_rescale(s_cmat[c], q);
            s_cpre = q;
        }
        __syncthreads();

        // meiosis_sampler.cc:155-187 for one locus, both possible values of the next locus's indicator:
        // bit 0 = image of 0, bit 1 = image of 1
        auto map_code = [&](int i, double f0, double f1) -> uint32_t {
            const double u = s_u[step * cap + i - cta_lo];
            if(i == M - 1) { const int x = ms_pick(f0, f1, u); return (uint32_t)(x | (x << 1)); }
            const double th = (i + 1 < cta_hi) ? s_theta[i + 1 - cta_lo] : theta_hi, ith = 1.0 - th;
            const int x0 = ms_pick(f0 * ith, f1 * th, u);     // next = 0: j != next ? theta : 1 - theta
            const int x1 = ms_pick(f0 * th, f1 * ith, u);     // next = 1
            return (uint32_t)(x0 | (x1 << 1));
        };

        uint32_t gmap = 2u;                              // identity
        {
            // exclusive prefix of this thread = (lane before) x (warps before) x (CTAs before)
            Mat2 q = mat2_shfl_up(acc, 1);
            if(lane == 0) { q.a = 1.0; q.b = 0.0; q.c = 0.0; q.d = 1.0; }
            if(warp > 0) q = mat2_mul_rescale(q, s_wmat[warp - 1]);
            if(rank > 0) q = mat2_mul_rescale(q, s_cpre);
            if(lo < hi) {
                double2 r = lo == 0 ? s_raw[0] : make_double2(0.0, 0.0);
                double v0, v1;
                int i = lo;
                if(lo == 0) {
                    const double tot0 = r.x + r.y;
                    v0 = r.x / tot0; v1 = r.y / tot0;        // fb_matrix[0..1] (meiosis_sampler.cc:134-137)
                    if(step == 0) { L.fb[0] = v0; L.fb[1] = v1; }
                    s_map[0] = (uint8_t) map_code(0, v0, v1);
                    i = 1;
                }
                else {
                    const double2 r0 = s_r0;
                    const double tot0 = r0.x + r0.y;
                    const double u0 = r0.x / tot0, u1 = r0.y / tot0;
                    const double w0 = q.a * u0 + q.b * u1, w1 = q.c * u0 + q.d * u1, tot = w0 + w1;
                    v0 = w0 / tot; v1 = w1 / tot;
                }
                for(; i < hi; ++i) {                       // the reference's recurrence (:140-153), same operation order
                    r = s_raw[i - cta_lo];
                    const double th = s_theta[i - cta_lo], ith = 1.0 - th;
                    double f0 = r.x * ((v1 * th) + (v0 * ith));
                    double f1 = r.y * ((v0 * th) + (v1 * ith));
                    const double inv = 1.0 / (f0 + f1);
                    f0 *= inv; f1 *= inv;
                    if(step == 0) { L.fb[2 * i] = f0; L.fb[2 * i + 1] = f1; }
                    s_map[i - cta_lo] = (uint8_t) map_code(i, f0, f1);
                    v0 = f0; v1 = f1;
                }
                for(i = hi - 1; i >= lo; --i) gmap = compose(s_map[i - cta_lo], gmap);
            }
        }

        if(tr) tr[5 + 8 * step] = clock64();
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
        if(tr) tr[6 + 8 * step] = clock64();
        if(t == 0) {
            uint32_t in = 2u;
            for(int c = SLK_MS_CLUSTER - 1; c > rank; --c) in = compose(s_cmap[c], in);
            s_cpost = in;
        }
        __syncthreads();
        {
            // the map from "right of everything" to the value entering this chunk from the right; the last
            // locus's map is constant, so the composition is too and may be evaluated at 0
            uint32_t in = __shfl_down_sync(0xffffffffu, gmap, 1);
            if(lane == 31) in = 2u;
            if(warp + 1 < nwarps) in = compose(in, s_wmap[warp + 1]);
            in = compose(in, s_cpost);
            if(lo < hi) {
                uint32_t x = apply(in, 0);
                const uint32_t pbit = 1u << L.step_parent[step];
                for(int i = hi - 1; i >= lo; --i) {
                    const int k = i - cta_lo;
                    x = apply(s_map[k], x);
                    const uint32_t bit = bits[i] & 1u;
                    if(bit != x) {
                        uint8_t* cell = L.dgp + (size_t) i * N + L.step_person[step];
                        *cell = (uint8_t)(*cell ^ pbit);
                        s_lncur[k] = s_lnflip[k];            // carried likelihood of the (new) current graph
                        if(step == 0) s_changed[k] = 1;
                        // the typed people's labels below this meiosis are now out of date at this locus
                        const int P = L.step_person[step];
                        for(int w = 0; w < W; ++w) L.stale[(size_t) i * W + w] |= L.ms.desc_mask[(size_t)(P - L.F) * W + w];
                        const int k0 = L.ms.typed_index[P];
                        if(k0 >= 0) { const uint32_t q = 2 * k0 + L.step_parent[step]; L.stale[(size_t) i * W + (q >> 5)] |= 1u << (q & 31); }
                    }
                }
            }
        }
        if(tr) tr[7 + 8 * step] = clock64();
        // The CTA's own tables are reused by the next step: a CTA barrier.  The tables other CTAs write into
        // (s_cmat, s_r0: before the next step's first cluster barrier; s_cmap: before its second) need none: a CTA can
        // only get that far after every CTA has arrived at the barrier that follows this step's last read of them.
        __syncthreads();
        if(tr) tr[8 + 8 * step] = clock64();
    }
    for(int i = cta_lo + t; i < cta_hi; i += T) L.lncur[i] = s_lncur[i - cta_lo];
    MS_TL(3);
}

// ---- descent-graph likelihood (descent_graph.cc:150-265) ------------------------------------------

// out[0] = sum over loci of ln L (or -DBL_MAX if any is impossible), out[1] = sum over intervals of the
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
        const double v = L.lncur[l];
        if(!(v > -DBL_MAX)) s_zero = 1; else a += v;
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
