// slk_types.h -- internal layout of the flattened, device-resident peel plan.
//
// The reference keeps one heap object per (op, locus) with pointer-chased members
// (gpu_lodscores.cc:396-507: ~7M cudaMallocs at 200 ops x 10k loci).  Here the whole plan is a
// few flat arrays, independent of the number of loci:
//
//   program stream   one variable-length record of 32-bit words per peel op (layout below),
//                    staged once per CTA into shared memory;
//   lops / lpf       the forward schedule: ops grouped into dependency levels; inside a level the
//                    ROWS of all ops (a row = one valid assignment of cutset digits 1..c-1; the
//                    thread that owns it evaluates the 4 values of digit 0 x 4 peel genotypes)
//                    form one index space, lpf = exclusive prefix of rows per op, so every thread
//                    of the team finds its (op, row) by a short binary search -- a level of 60
//                    small ops costs one pass, not 60 warp-items;
//   bops             the backward (sampling) schedule: ops grouped into levels by "who peels
//                    my cutset", four lanes per op;
//   arena            one slab of doubles per team holding every live peel matrix; offsets
//                    below `smem_doubles` live in shared memory, the rest in an L2-resident
//                    global scratch slab (only needed when sum 4^c outgrows 227 KB).
//
// Two programs are built from the same peel sequence: the sampler program (all matrices live
// until the backward pass) and the trait program (matrices freed after their single consumer,
// peeling.h:136-143 / peel_sequence_generator.cc:70-82 guarantee exactly one).
#ifndef SLK_TYPES_H
#define SLK_TYPES_H

#include <stdint.h>

#define SLK_SLOTS      4            // cells per thread per quad (16 independent product chains)
#define SLK_POS_PEEL   15           // digit position meaning "the peel node itself"
#define SLK_NO_SHIFT   0xFF

// ---- program stream record -------------------------------------------------------------
//  w0  type[0:3) | c[4:8) | nprev[8:12) | nkids[12:16) | peelnode[16:32)
//  w1  arena offset of this op's matrix (doubles)
//  w2  dense offset (sum of 4^c of earlier ops, peel order) -- parity dumps only
//  w3  flags: bit0 = some previous function is keyed on the peel node
//  w4,w5  static legal masks, 4 bits per cutset digit (trait program: disease_prob != 0;
//         sampler program: unused, the per-locus masks come from the elimination table)
//  w6  static legal mask of the peel node (low 4 bits)
//  then ceil(c/2) words: cutset person ids, two 16-bit ids per word
//  then per previous function:
//        word A  arena offset of the consumed matrix
//        word B  peel_shift[0:8) (SLK_NO_SHIFT if the peel node is not one of its keys) | nruns[8:16)
//                | d0_shift[16:24) (where the consumer's cutset digit 0 sits in it, or SLK_NO_SHIFT)
//        ceil(nruns/2) words of 16-bit runs over the consumer's digits 1..c-1:
//                src_shift[0:5) | dst_shift[5:10) | nbits[10:15)
//  then per child (CHILD_PEEL: the peel node itself; PARENT_PEEL: cutset members that are its
//  offspring, cutset order):
//        person[0:16) | kid_pos[16:20) | mat_pos[20:24) | pat_pos[24:28) | male[28]
#define SLK_REC_HEADER 7

struct SlkProgram {
    const uint32_t* stream;         // [stream_words]
    const uint16_t* op_start;       // [nops] word offset of each record
    const uint16_t* lops;           // [nops] ops in forward-level order (largest first inside a level)
    const uint32_t* lpf;            // [nops] quads before this op inside its level
    const uint32_t* flevel_quads;   // [n_flevels] quads of each level
    const uint16_t* flevel_start;   // [n_flevels + 1] into lops
    const uint16_t* bops;           // ops in backward-level order
    const uint16_t* blevel_start;   // [n_blevels + 1] into bops
    int stream_words;
    int n_flevels;
    int n_blevels;
    int arena_doubles;              // whole arena
    int smem_doubles;               // arena prefix kept in shared memory
    int team_threads;               // 32, 64, 128, 256 or 512
    int cta_threads;
    int prog_smem_bytes;            // CTA-shared staging of stream/op_start/items/levels
    int team_smem_bytes;            // per team: arena prefix + tables
    int table_doubles_per_child;    // 16 (sampler) or 4 (trait)
};

// M-sampler tables (slk_msampler.cuh)
struct SlkMsPlan {
    const uint16_t* seq;            // [N-F] non-founders, parents before children (meiosis_sampler.cc:41-72)
    const uint16_t* typed;          // [n_typed] Person::istyped(), pedigree order
    const uint8_t* obsT;            // [n_typed][M] observed genotype (SLK_UNTYPED..SLK_HOMOZ_B)
    const double* minor;            // [M] Snp::minor()
    const double* lnmajor;          // [M] ln(1 - minor) (-1e300 if the frequency is 0)
    const double* lnminor;          // [M] ln(minor)
    const uint32_t* desc_mask;      // [N-F][W] te slots (2k + parent) whose lineage can pass through that person
    const int16_t* typed_index;     // [N] index into typed[] or -1
    int W;                          // words per slot mask = ceil(2 n_typed / 32)
    int n_typed;
    int available;                  // slk_problem.minor_freq was given
};

struct SlkDevPlan {
    int N, F, M, nlod, sex_linked, nops;
    int last_op;                    // op whose 1-cell matrix is the likelihood
    const int16_t* mother;          // [N] (-1 founders)
    const int16_t* father;          // [N]
    const uint8_t* male;            // [N]
    const double* disease_prob;     // [N][4]
    const uint8_t* gcode;           // [M][N]: legal mask (trait encoding) | prior class << 4
    const double* fprior;           // [M][2][4] founder priors (autosomal/female, X male)
    const double* theta;            // [M-1]
    const double* partial;          // [M-1]
    const double* log_theta;        // [M-1] ln theta
    const double* log_1mtheta;      // [M-1] ln (1 - theta)
    double marker_transmission;     // descent_graph.cc:22,35
    SlkProgram ls;                  // sampler program
    SlkProgram lod;                 // trait program
    SlkMsPlan ms;                   // M-sampler tables
};

// Per-team shared-memory layout (byte offsets from the team's base), used by the host to size
// launches and by the kernels to carve the slab.
struct SlkTeamLayout {
    int arena;      // smem_doubles doubles
    int tables;     // (N-F)*k doubles: sampler transmission tables (k = 16) / trait recombination weights (k = 4)
    int scal;       // 48 doubles: thetas [0..4), class priors [16..36), founder priors of the locus [36..44)
    int lmw;        // nops x uint64 per-locus legal masks of each op's cutset (sampler)
    int ru;         // nops doubles: the genotype draws of the current locus (sampler)
    int bytes;      // 4 byte arrays of round16(N): gcode row, dg left, dg right, sampled genotypes
    int nbytes;     // round16(N)
    int red;        // 32 ints of reduction scratch
    int total;
};

#if defined(__CUDACC__)
__host__ __device__
#endif
static inline SlkTeamLayout slk_team_layout(int N, int F, int nops, int smem_doubles, int table_doubles_per_child) {
    SlkTeamLayout L;
    L.arena = 0;
    L.tables = (smem_doubles * 8 + 15) & ~15;
    L.scal = L.tables + (N - F) * table_doubles_per_child * 8;
    L.lmw = L.scal + 48 * 8;
    L.ru = L.lmw + nops * 8;
    L.bytes = L.ru + nops * 8;
    L.nbytes = (N + 15) & ~15;
    L.red = L.bytes + 4 * L.nbytes;
    L.total = (L.red + 32 * 4 + 15) & ~15;
    return L;
}

// prior classes (person.cc:224-299 collapses to these; see slk_plan.cc)
enum {
    SLK_PRIOR_UU = 0,        // (1,0,0,0)
    SLK_PRIOR_AA = 1,        // (0,1,0,0)
    SLK_PRIOR_HET = 2,       // (0,0,.5,.5)
    SLK_PRIOR_FLAT = 3,      // (.25,.25,.25,.25)
    SLK_PRIOR_XMALE = 4,     // (.5,.5,0,0)
    SLK_PRIOR_FOUNDER = 5,   // fprior[locus][0]
    SLK_PRIOR_FOUNDER_X = 6, // fprior[locus][1]
    SLK_PRIOR_PERSON = 7     // disease_prob[person] (ELOD's simulated trait locus, person.h:204-208)
};

#endif
