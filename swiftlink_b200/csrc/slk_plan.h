// slk_plan.h -- host-side flattening of a peel sequence into the device plan (slk_types.h).
#ifndef SLK_PLAN_H
#define SLK_PLAN_H

#include <stdint.h>
#include <string>
#include <vector>

#include "swiftlink_b200.h"
#include "slk_types.h"

namespace slk {

struct HostProgram {
    std::vector<uint32_t> stream;
    std::vector<uint16_t> op_start;  // record offset / 4
    std::vector<uint16_t> imap;      // record offset / 4 of the op of every group of four forward rows
    std::vector<uint32_t> dbg;       // [nops][3] dense offset, refpos (parity dumps)
    std::vector<uint32_t> flevel_items;
    std::vector<uint32_t> flevel_map;
    std::vector<uint16_t> bops;
    std::vector<uint16_t> blevel_start;
    std::vector<int> mat_off;        // arena offset per op (doubles)
    std::vector<uint8_t> mat_pad;    // 1: padded layout (two doubles after every 16)
    std::vector<int> flevel;         // forward level of each op
    std::vector<uint8_t> blob;       // the device image of the program (slk_types.h)
    SlkProgram layout;               // offsets into the blob + geometry (blob pointer unset)
    int arena_doubles;
    int smem_doubles;
    int team_threads;
    int cta_threads;
    int prog_smem_bytes;
    int team_smem_bytes;
    int table_doubles_per_child;
    long valid_cells;                // static valid cells (trait) or dense cells (sampler)
};

struct HostPlan {
    int N, F, M, nlod, sex_linked, nops, last_op;
    std::vector<int16_t> mother, father;
    std::vector<uint8_t> male;
    std::vector<double> disease_prob;    // [N][4]
    std::vector<double> person_prior;    // [N][4] prior of the SLK_PRIOR_PERSON class (ELOD's simulated trait locus)
    std::vector<uint8_t> gcode;          // [M][N]
    std::vector<double> fprior;          // [M][2][4]
    std::vector<double> theta, partial, log_theta, log_1mtheta;
    double marker_transmission;
    std::vector<int> dense_off;          // per op
    long sum_cells, sum_presum;          // sum 4^c, sum 4^(c+1)
    double flops_ls, flops_lod;          // SURVEY.md section 8(d) F_L and F_T
    int max_cutset;
    HostProgram ls, lod;
    // M-sampler
    bool ms_available;
    std::vector<uint16_t> ms_seq, ms_typed;
    std::vector<uint8_t> ms_typed_auto;  // [nt] x-linked male: the maternal allele only
    std::vector<uint8_t> ms_obsT;        // [n_typed][M]
    std::vector<double> ms_minor, ms_lnmajor, ms_lnminor;   // [M]
    std::vector<int32_t> ms_ordering;    // meioses an M-sweep visits (markov_chain.cc:68-80)
    std::vector<uint32_t> ms_desc_mask;  // [N-F][W]
    std::vector<int16_t> ms_typed_index; // [N]
    int ms_W;
};

// shared-memory budget per SM the layout may assume (bytes); 227 KB on sm_100
static const int kSmemPerBlockMax = 232448;

// Builds the flattened plan.  Returns false and fills err on a malformed problem.
bool build_plan(const slk_problem& pb, HostPlan& out, std::string& err);

}  // namespace slk

#endif
