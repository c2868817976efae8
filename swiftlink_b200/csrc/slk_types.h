// slk_types.h -- internal layout of the flattened, device-resident peel plan.
//
// The reference keeps one heap object per (op, locus) with pointer-chased members
// (gpu_lodscores.cc:396-507: ~7M cudaMallocs at 200 ops x 10k loci).  Here the whole plan is a
// few flat arrays, independent of the number of loci:
//
//   program stream   one record of 32-bit words per peel op (layout below), every field the kernels
//                    need already decoded by the host, staged once per CTA into shared memory;
//   imap             the forward schedule: ops grouped into dependency levels; inside a level the
//                    ROWS of all ops (a row = one assignment of cutset digits 1..c-1; the thread that
//                    owns it evaluates the 4 values of digit 0 x 4 peel genotypes = 16 product chains)
//                    form one index space, every op's rows padded to a multiple of four, and
//                    imap[item >> 2] names the op -- one shared-memory load instead of a search;
//   bops             the backward (sampling) schedule: ops grouped into levels by "who peels
//                    my cutset", four lanes per op;
//   arena            one slab of doubles per team holding every live peel matrix; offsets
//                    below `smem_doubles` live in shared memory, the rest in an L2-resident
//                    global scratch slab (only needed when sum 4^c outgrows 227 KB).
//
// MATRIX LAYOUT.  The reference indexes a peel matrix by its cutset in the order the
// PeelSequenceGenerator happened to find it (peel_matrix.h:37-45).  Here every matrix is indexed
// by its cutset SORTED BY PEEL POSITION (the member that is peeled soonest is digit 0).  Every
// function a peel op consumes is keyed on a subset of (cutset + peel node), so under one global
// order of people its digits are a SUBSEQUENCE of the consumer's digits: the peel node -- peeled
// now, hence first -- is its digit 0 when present, the consumer's digit 0 comes next.  A thread's
// 4 x 4 tile therefore reads 4 or 16 CONSECUTIVE doubles of each consumed matrix (vector loads,
// whole 32-byte sectors from the L2 slab) and writes 4 consecutive doubles; the lanes of a warp,
// which differ in the row, touch consecutive blocks.  Matrices that live in shared memory insert
// two doubles of padding after every 16 (SLK_PAD) so that those blocks fall into different banks.
// The reference's cell order only matters in the parity dumps (digit permutation `refpos`).
//
// Two programs are built from the same peel sequence: the sampler program (all matrices live
// until the backward pass) and the trait program (matrices freed after their single consumer,
// peeling.h:136-143 / peel_sequence_generator.cc:70-82 guarantee exactly one).
#ifndef SLK_TYPES_H
#define SLK_TYPES_H

#include <stdint.h>

#define SLK_POS_PEEL   15           // digit position meaning "the peel node itself"

// ---- program stream record (all offsets in 32-bit words; records start on 16-byte boundaries) ----
//  0  type[0:3) | c[4:8) | nprev[8:12) | nkids[12:16) | peelnode[16:32)
//  1  arena offset of this op's matrix (doubles) | SLK_MAT_PAD if the matrix uses the padded layout
//  2  first item of this op inside its forward level
//  3  bit0 = some previous function is keyed on the peel node | static legal mask of the peel node [8:12)
//     (trait program: disease_prob != 0) | op id [16:32)
//  then ceil(c/2) words: the sorted cutset, two 16-bit person ids per word
//  then (trait program only) 2 words: static legal masks, 4 bits per sorted digit
//  then per previous function, in the reference's order (4 words):
//        +0  arena offset of the consumed matrix | SLK_MAT_PAD
//        +1  kind[0:8) (SLK_KEY_*) | nruns[8:16) | run 4 [16:32)
//        +2  runs 0, 1   16-bit runs gathering the consumed matrix's ROW index from the row digits of the
//        +3  runs 2, 3   consumer's cell: src_bit[0:5) | dst_bit[5:10) | nbits[10:15)
//  then per child (CHILD_PEEL: the peel node itself; PARENT_PEEL: cutset members that are its
//  offspring, in the reference's children order), one word:
//        person[0:16) | kid_pos[16:20) | mat_pos[20:24) | pat_pos[24:28) | male[28]   (positions: sorted digits)
// The parity dumps need two more facts per op, kept out of the hot record in SlkDevPlan::dbg: the dense offset
// (sum of 4^c of earlier ops, peel order) and `refpos`, 4 bits per sorted digit = its position in the reference's
// cutset order.
#define SLK_REC_HEADER 4
#define SLK_REC_PREV   4
#define SLK_LEVEL_FINE 0x80000000u // forward level with few rows: one slot per thread instead of a whole 4 x 4 tile
#define SLK_MAT_PAD    0x80000000u
#define SLK_OFF_MASK   0x7fffffffu

// how a consumed matrix is keyed relative to the consumer's tile (sorted layout: four cases only)
enum {
    SLK_KEY_VS = 0,     // digit 0 = consumer's peel node, digit 1 = consumer's digit 0: index = v + 4 s + 16 R
    SLK_KEY_V  = 1,     // digit 0 = consumer's peel node:                               index = v + 4 R
    SLK_KEY_S  = 2,     // digit 0 = consumer's digit 0:                                 index = s + 4 R
    SLK_KEY_R  = 3      // keyed on row digits only:                                     index = R
};

// per (op, locus) record rebuilt at the start of every unit (sampler) / once per launch (trait): 16 bytes
//   lmw   legal-genotype masks of the sorted cutset, 4 bits per digit
//   nrows valid rows (product of popcounts of digits 1..c-1)
//   misc  legal mask of digit 0 [0:4) (1 if c == 0) | legal mask of the peel node [4:8)
struct SlkOpLoc {
    unsigned long long lmw;
    uint32_t nrows;
    uint32_t misc;
};

// The program is ONE blob of bytes on the device, copied verbatim into shared memory by every CTA
// (a single 1-D bulk copy); the arrays sit at 16-byte aligned offsets:
struct SlkProgram {
    const uint8_t* blob;            // [blob_bytes]
    int blob_bytes;                 // multiple of 16
    int off_stream;                 // uint32[stream_words]
    int off_op_start;               // uint16[nops] record offset / 4
    int off_imap;                   // uint16[imap_len] record offset / 4 of the op of every group of four forward items, level after level
    int off_flevel_items;           // uint32[n_flevels] items of each level
    int off_flevel_map;             // uint32[n_flevels] first imap entry of each level | SLK_LEVEL_FINE
    int trait;                      // 1: trait program (records carry the static legal masks)
    int off_bops;                   // uint16[nops] ops in backward-level order
    int off_blevel_start;           // uint16[n_blevels + 1] into bops
    int off_glist;                  // uint8[16] genotype lists of the 16 legal masks
    int off_dprob;                  // double[N][4] disease probabilities (trait program only, else -1)
    int stream_words;
    int imap_len;
    int n_flevels;
    int n_blevels;
    int arena_doubles;              // whole arena
    int smem_doubles;               // arena prefix kept in shared memory
    int team_threads;               // 32, 64, 128, 256 or 512
    int cta_threads;
    int prog_smem_bytes;            // blob_bytes + 16: CTA-shared copy of the program and the mbarrier of its bulk copy
    int team_smem_bytes;            // per team: arena prefix + tables
    int table_doubles_per_child;    // 8 (sampler) or 4 (trait)
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
    const double* person_prior;     // [N][4] prior of the SLK_PRIOR_PERSON class
    const uint8_t* gcode;           // [M][N]: legal mask (trait encoding) | prior class << 4
    const double* fprior;           // [M][2][4] founder priors (autosomal/female, X male)
    const double* theta;            // [M-1]
    const double* partial;          // [M-1]
    const double* log_theta;        // [M-1] ln theta
    const double* log_1mtheta;      // [M-1] ln (1 - theta)
    const uint32_t* dbg;            // [nops][3] dense offset, refpos (parity dumps only)
    double marker_transmission;     // descent_graph.cc:22,35
    SlkProgram ls;                  // sampler program
    SlkProgram lod;                 // trait program
    SlkMsPlan ms;                   // M-sampler tables
};

// Per-team shared-memory layout (byte offsets from the team's base), used by the host to size
// launches and by the kernels to carve the slab.
struct SlkTeamLayout {
    int arena;      // smem_doubles doubles
    int tables;     // (N-F)*k doubles: sampler transmission tables (k = 8) / trait recombination weights (k = 4)
    int scal;       // 48 doubles: thetas [0..4), class priors [16..36), founder priors of the locus [36..44)
    int oploc;      // nops x SlkOpLoc (16 bytes)
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
    L.oploc = L.scal + 48 * 8;
    L.ru = L.oploc + nops * 16;
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
