// slk_plan.cc -- flattens PeelSequenceGenerator's output into the device plan.
//
// Reference semantics reproduced here (all under /root/reference/src):
//   * index digit i of a peel matrix is cutset[i]; the peel node is the top digit of the
//     presum matrix                                 (peel_sequence_generator.cc:95-104)
//   * a previous function is looked up by gathering ITS cutset digits from the consumer's
//     (cutset + peel node) assignment               (peel_matrix.h:37-45, rfunction.cc:132-134)
//   * sampler validity = GenotypeElimination::is_legal on every cutset digit
//                                                   (peel_sequence_generator.cc:109-159)
//   * trait validity   = disease_prob != 0 on every cutset digit      (:162-179)
//   * marker priors    = Person::populate_trait_prob_cache            (person.cc:224-299)
#include "slk_plan.h"
#include "slk_peel.h"
#include "slk_geometry.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <tuple>

namespace slk {

namespace {

int find_pos(const slk_peel_op& op, int person) {
    for(int i = 0; i < op.ncut; ++i) if(op.cutset[i] == person) return i;
    if(op.peelnode == person) return SLK_POS_PEEL;
    return -1;
}

// genotype.cc:91-102 turned around: legal-mask bit v (trait code) from the elimination mask
uint8_t legal_mask_from_elimination(uint8_t m) {
    uint8_t r = 0;
    if(m & 8) r |= 1 << SLK_UU;     // AA
    if(m & 1) r |= 1 << SLK_AA;     // BB
    if(m & 2) r |= 1 << SLK_AU;     // BA
    if(m & 4) r |= 1 << SLK_UA;     // AB
    return r;
}

// person.cc:247-299 collapsed to a class; the arithmetic that remains locus dependent
// (founder with no genotype constraint) is done once per locus in founder_prior()
int prior_class(bool founder, bool typed, int g, bool xmale) {
    if(typed) {
        switch(g) {
            case SLK_HETERO:  return SLK_PRIOR_HET;
            case SLK_HOMOZ_A: return SLK_PRIOR_UU;
            case SLK_HOMOZ_B: return SLK_PRIOR_AA;
            default:
                // typed person, untyped marker: 1.0 / marker_prob for every genotype -- the
                // reference does not apply the X-male restriction on this branch
                if(!founder) return SLK_PRIOR_FLAT;
                return xmale ? SLK_PRIOR_FOUNDER_X : SLK_PRIOR_FOUNDER;
        }
    }
    if(!founder) return xmale ? SLK_PRIOR_XMALE : SLK_PRIOR_FLAT;
    return xmale ? SLK_PRIOR_FOUNDER_X : SLK_PRIOR_FOUNDER;
}

// person.cc:224-245 for an unconstrained founder: probs[j] = map prob (hetero zeroed for an
// untyped X male), then divided by their sum
void founder_prior(const double* mapprob, bool zero_hetero, double* out) {
    double p[4];
    for(int j = 0; j < 4; ++j) p[j] = mapprob[j];
    if(zero_hetero) { p[SLK_AU] = 0.0; p[SLK_UA] = 0.0; }
    double total = p[0] + p[1] + p[2] + p[3];
    for(int j = 0; j < 4; ++j) out[j] = p[j] / total;
}

struct Liveness {
    // first-fit allocator over [offset, offset+size) intervals
    std::vector<std::pair<int,int> > free_list;   // (offset, size), sorted by offset
    int top;
    Liveness() : top(0) {}
    int alloc(int size) {
        for(size_t i = 0; i < free_list.size(); ++i) {
            if(free_list[i].second >= size) {
                int off = free_list[i].first;
                free_list[i].first += size;
                free_list[i].second -= size;
                if(free_list[i].second == 0) free_list.erase(free_list.begin() + i);
                return off;
            }
        }
        int off = top;
        top += size;
        return off;
    }
    void release(int off, int size) {
        free_list.push_back(std::make_pair(off, size));
        std::sort(free_list.begin(), free_list.end());
        for(size_t i = 0; i + 1 < free_list.size(); ) {
            if(free_list[i].first + free_list[i].second == free_list[i+1].first) {
                free_list[i].second += free_list[i+1].second;
                free_list.erase(free_list.begin() + i + 1);
            }
            else ++i;
        }
        if(!free_list.empty()) {
            std::pair<int,int>& last = free_list.back();
            if(last.first + last.second == top) { top = last.first; free_list.pop_back(); }
        }
    }
};

int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

// padded size of a 4^c matrix: two doubles after every 16 (shared-memory bank spreading), even
static int padded_size(int ncut) {
    const int n = 1 << (2 * ncut);
    const int p = n + ((n >> 4) << 1);
    return (p + 1) & ~1;
}

static bool emit_program(const slk_problem& pb, const HostPlan& hp, bool trait, HostProgram& pr, std::string& err) {
    const int nops = pb.n_ops;
    const int N = pb.n_members;

    // ---- the global order of people: position in the peel sequence ----
    std::vector<int> ppos(N, -1);
    for(int i = 0; i < nops; ++i) ppos[pb.ops[i].peelnode] = i;
    std::vector<std::vector<int> > scut(nops);
    for(int i = 0; i < nops; ++i) {
        const slk_peel_op& op = pb.ops[i];
        scut[i].assign(op.cutset, op.cutset + op.ncut);
        std::sort(scut[i].begin(), scut[i].end(), [&](int a, int b) { return ppos[a] < ppos[b]; });
        for(int d = 0; d < op.ncut; ++d) {
            if(ppos[scut[i][d]] <= i) {
                std::ostringstream e; e << "op " << i << ": cutset member " << scut[i][d] << " is peeled before the op that is keyed on it";
                err = e.str(); return false;
            }
        }
    }
    auto spos = [&](int i, int person) -> int {
        for(size_t d = 0; d < scut[i].size(); ++d) if(scut[i][d] == person) return (int) d;
        if(pb.ops[i].peelnode == person) return SLK_POS_PEEL;
        return -1;
    };

    // ---- forward levels: an op runs after every function it consumes ----
    std::vector<int> flevel(nops, 0), consumer(nops, -1);
    int n_flevels = 0;
    for(int i = 0; i < nops; ++i) {
        int lv = 0;
        for(int j = 0; j < pb.ops[i].nprev; ++j) {
            lv = std::max(lv, flevel[pb.ops[i].prev[j]] + 1);
            consumer[pb.ops[i].prev[j]] = i;
        }
        flevel[i] = lv;
        n_flevels = std::max(n_flevels, lv + 1);
    }
    pr.flevel = flevel;

    // ---- backward levels: an op is sampled after the ops that peel its cutset members ----
    std::vector<int> blevel(nops, 0);
    int n_blevels = 0;
    for(int i = nops - 1; i >= 0; --i) {
        int lv = 0;
        for(int d = 0; d < pb.ops[i].ncut; ++d) lv = std::max(lv, blevel[ppos[pb.ops[i].cutset[d]]] + 1);
        blevel[i] = lv;
        n_blevels = std::max(n_blevels, lv + 1);
    }

    // ---- static validity: cells per op (sorted digits) ----
    std::vector<long> ncells(nops), nrows(nops);
    std::vector<uint64_t> static_lmw(nops, 0);
    std::vector<uint8_t> dmask(N, 0);
    for(int i = 0; i < N; ++i)
        for(int g = 0; g < 4; ++g) if(pb.disease_prob[i*4 + g] != 0.0) dmask[i] |= (uint8_t)(1 << g);
    pr.valid_cells = 0;
    for(int i = 0; i < nops; ++i) {
        const slk_peel_op& op = pb.ops[i];
        long n = 1, rows = 1;
        for(int d = 0; d < op.ncut; ++d) {
            uint8_t m = trait ? dmask[scut[i][d]] : 15;
            static_lmw[i] |= (uint64_t) m << (4 * d);
            n *= __builtin_popcount(m);
            if(d > 0) rows *= __builtin_popcount(m);
        }
        ncells[i] = n;
        // rows = valid assignments of digits 1..c-1 (digit 0 is the register tile's slot axis)
        nrows[i] = rows;
        pr.valid_cells += n;
    }

    // ---- geometry (needed before the arena: what stays in shared memory depends on it) ----
    const int tpc_doubles = trait ? 4 : 8;
    pr.table_doubles_per_child = tpc_doubles;
    const int table_bytes = slk_team_layout(N, pb.n_founders, nops, 0, tpc_doubles).total;
    const long work = trait ? pr.valid_cells : hp.sum_cells;
    int team = 32;
    if(work > 1024)  team = 64;
    if(work > 4096)  team = 128;
    if(work > 8192)  team = 256;
    long padded_total = 0;
    for(int i = 0; i < nops; ++i) padded_total += padded_size(pb.ops[i].ncut);
    // One unit is latency bound (a chain of ~50 dependent levels), so the SM needs several units in
    // flight.  When the whole arena of a team would crowd the others out of shared memory, only the
    // small matrices (cutset <= 4, the ones hit by the many small latency-bound ops) stay in shared
    // memory and the few large ones go to the L2-resident global slab.
    const bool hybrid = (size_t) padded_total * 8 + table_bytes > 100 * 1024;
    // Large plans: one warp per team and as many teams as shared memory holds.  A unit is a chain of ~50 dependent
    // levels most of which hold a few dozen tiles, so what fills the SM is units in flight, not threads per unit; a
    // one-warp team also needs no block-level barrier.  Measured on B200, 200-member pedigree, ms per L-sweep:
    // 12 x 32 threads 3.05, 8 x 64 3.22, 6 x 64 3.58, 5 x 128 3.90, 3 x 128 4.38, 3 x 192 4.54, 3 x 256 5.23.
    if(hybrid) team = 32;
    // The trait program writes and reads every matrix once per unit, all of it in the global slab.  Six teams of three
    // warps per SM halve the slab of the resident teams (888 x 115 KB on the 200-member pedigree; ncu: 10.9 GB of DRAM
    // traffic per pass against 15.3 GB) and are faster than twelve one-warp teams.  Measured on B200, ms per scoring pass:
    // 6 x 96 threads 12.30, 6 x 64 12.45, 5 x 128 12.83, 8 x 64 13.06, 5 x 96 13.61, 4 x 128 14.28, 3 x 192 14.50,
    // 3 x 256 15.54, 12 x 32 15.94, 4 x 64 16.00.
    if(hybrid && trait) team = 96;
    team = env_int(trait ? "SLK_LOD_TEAM" : "SLK_LS_TEAM", team);
    if(slk_max_cta(team) == 0) team = 128;
    pr.team_threads = team;
    // the CTA holds a whole number of teams and at most the thread count the kernels of this team size are
    // compiled for (slk_geometry.h)
    const int max_cta = slk_max_cta(team);
    int cta = env_int(trait ? "SLK_LOD_CTA_THREADS" : "SLK_LS_CTA_THREADS", env_int("SLK_CTA_THREADS", hybrid ? max_cta : 128));
    cta = std::max(team, std::min(cta, max_cta));
    cta = (cta / team) * team;
    pr.cta_threads = cta;
    int teams_per_cta = pr.cta_threads / team;

    // ---- record offsets ----
    int stream_words = 0;
    pr.op_start.assign(nops, 0);
    for(int i = 0; i < nops; ++i) {
        const slk_peel_op& op = pb.ops[i];
        int nkids = (op.type == SLK_CHILD_PEEL) ? 1 : (op.type == SLK_PARENT_PEEL ? op.nchild : 0);
        if(stream_words / 4 > 65535) { err = "peel program too large for 16-bit record offsets"; return false; }
        pr.op_start[i] = (uint16_t)(stream_words / 4);
        stream_words += round_up(SLK_REC_HEADER + ((op.ncut + 1) >> 1) + (trait ? 2 : 0) + SLK_REC_PREV * op.nprev + nkids, 4);
    }

    // ---- forward schedule: ops per level, every op's rows padded to a multiple of four items ----
    pr.imap.clear(); pr.flevel_items.clear(); pr.flevel_map.clear();
    std::vector<uint32_t> qstart(nops, 0);
    for(int l = 0; l < n_flevels; ++l) {
        // big ops first inside a level, ops of the same shape next to each other so that the threads of a
        // warp that hold different ops still follow the same control flow
        std::vector<int> ops_here;
        for(int i = 0; i < nops; ++i) if(flevel[i] == l) ops_here.push_back(i);
        auto shape = [&](int i) {
            const slk_peel_op& o = pb.ops[i];
            return std::make_tuple(-nrows[i], o.type, o.nprev, o.nchild, o.ncut);
        };
        std::stable_sort(ops_here.begin(), ops_here.end(), [&](int a, int b) { return shape(a) < shape(b); });
        uint32_t items = 0;
        const uint32_t map0 = (uint32_t) pr.imap.size();
        for(size_t q = 0; q < ops_here.size(); ++q) {
            const int i = ops_here[q];
            qstart[i] = items;
            const uint32_t groups = (uint32_t)((nrows[i] + 3) / 4);
            for(uint32_t g = 0; g < groups; ++g) pr.imap.push_back(pr.op_start[i]);
            items += 4 * groups;
        }
        // a level with few rows is latency bound: one (row, slot) per thread spreads it over four times as
        // many threads (SLK_LEVEL_FINE); items then count slots
        const bool fine = items * 4u <= 2u * (uint32_t) team;
        pr.flevel_map.push_back(map0 | (fine ? SLK_LEVEL_FINE : 0u));
        pr.flevel_items.push_back(fine ? 4u * items : items);
    }
    if(pr.imap.empty()) pr.imap.push_back(0);

    pr.bops.clear();
    pr.blevel_start.assign(1, 0);
    for(int l = 0; l < n_blevels; ++l) {
        // four lanes per op: keep ops of the same shape together so a warp follows one control flow
        std::vector<int> here;
        for(int i = nops - 1; i >= 0; --i) if(blevel[i] == l) here.push_back(i);
        std::stable_sort(here.begin(), here.end(), [&](int a, int b) {
            const slk_peel_op& x = pb.ops[a]; const slk_peel_op& y = pb.ops[b];
            return std::make_tuple(x.type, x.nprev, x.nchild, x.ncut) < std::make_tuple(y.type, y.nprev, y.nchild, y.ncut);
        });
        for(size_t q = 0; q < here.size(); ++q) pr.bops.push_back((uint16_t) here[q]);
        pr.blevel_start.push_back((uint16_t) pr.bops.size());
    }

    // CTA-shared copy of the program blob: stream, op_start, imap, level tables, genotype-list LUT, (trait)
    // disease probabilities, each at a 16-byte aligned offset
    SlkProgram& lay = pr.layout;
    memset(&lay, 0, sizeof(lay));
    {
        int o = 0;
        lay.off_stream = o;        o += round_up(stream_words * 4, 16);
        lay.off_op_start = o;      o += round_up(nops * 2, 16);
        lay.off_imap = o;          o += round_up((int) pr.imap.size() * 2, 16);
        lay.off_flevel_items = o;  o += round_up(n_flevels * 4, 16);
        lay.off_flevel_map = o;    o += round_up(n_flevels * 4, 16);
        lay.off_bops = o;          o += round_up(nops * 2, 16);
        lay.off_blevel_start = o;  o += round_up((n_blevels + 1) * 2, 16);
        lay.off_glist = o;         o += 16;
        lay.off_dprob = -1;
        if(trait) { lay.off_dprob = o; o += N * 32; }
        lay.blob_bytes = o;
    }
    pr.prog_smem_bytes = lay.blob_bytes + 16;          // + the mbarrier of the bulk copy that stages the blob

    // ---- arena ----
    pr.mat_off.assign(nops, 0);
    pr.mat_pad.assign(nops, 0);
    const int budget = kSmemPerBlockMax - pr.prog_smem_bytes;           // one CTA per SM worst case
    // as many teams as the CTA's shared memory holds (tables alone; the arena prefix takes what is left)
    while(teams_per_cta > 1 && budget / teams_per_cta < table_bytes) --teams_per_cta;
    pr.cta_threads = teams_per_cta * team;
    int per_team_budget = budget / teams_per_cta - table_bytes;
    if(per_team_budget < 0) per_team_budget = 0;
    int limit_doubles = per_team_budget / 8;
    limit_doubles = std::min(limit_doubles, std::max(0, env_int(trait ? "SLK_LOD_SMEM_DOUBLES" : "SLK_LS_SMEM_DOUBLES", limit_doubles)));

    if(trait) {
        // matrices die after the level of their single consumer; every matrix uses the padded layout
        std::vector<int> level_order;
        for(int lv = 0; lv < n_flevels; ++lv) for(int i = 0; i < nops; ++i) if(flevel[i] == lv) level_order.push_back(i);
        Liveness lv;
        std::vector<std::vector<int> > dies_after(n_flevels);
        for(int i = 0; i < nops; ++i) if(consumer[i] >= 0) dies_after[flevel[consumer[i]]].push_back(i);
        size_t k = 0;
        for(int l = 0; l < n_flevels; ++l) {
            for(; k < level_order.size() && flevel[level_order[k]] == l; ++k) {
                int i = level_order[k];
                pr.mat_off[i] = lv.alloc(padded_size(pb.ops[i].ncut));
                pr.mat_pad[i] = 1;
            }
            for(size_t q = 0; q < dies_after[l].size(); ++q) {
                int i = dies_after[l][q];
                lv.release(pr.mat_off[i], padded_size(pb.ops[i].ncut));
            }
        }
        int hw = 0;
        for(int i = 0; i < nops; ++i) hw = std::max(hw, pr.mat_off[i] + padded_size(pb.ops[i].ncut));
        pr.arena_doubles = hw;
        int smem_doubles = std::min(pr.arena_doubles, limit_doubles);
        if(smem_doubles < pr.arena_doubles) {
            // only whole matrices live in shared memory: cut at a matrix boundary no span straddles
            int cut = 0;
            std::vector<std::pair<int,int> > spans;
            for(int i = 0; i < nops; ++i) spans.push_back(std::make_pair(pr.mat_off[i], padded_size(pb.ops[i].ncut)));
            std::sort(spans.begin(), spans.end());
            for(size_t q = 0; q < spans.size(); ++q)
                if(spans[q].first + spans[q].second <= smem_doubles) cut = std::max(cut, spans[q].first + spans[q].second);
            bool ok = false;
            while(!ok) {
                ok = true;
                for(size_t q = 0; q < spans.size(); ++q)
                    if(spans[q].first < cut && spans[q].first + spans[q].second > cut) { cut = spans[q].first; ok = false; }
            }
            smem_doubles = cut;
        }
        pr.smem_doubles = smem_doubles;
    }
    else {
        // everything stays live until the backward pass; small matrices first: they form the shared-memory
        // prefix (padded layout), the large ones go to the global slab (plain layout, 128-byte aligned)
        std::vector<int> order(nops);
        for(int i = 0; i < nops; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return pb.ops[a].ncut < pb.ops[b].ncut; });
        const int maxcut_smem = env_int("SLK_LS_SMEM_MAXCUT", hybrid ? 4 : SLK_MAX_CUTSET);
        int off = 0;
        size_t q = 0;
        for(; q < order.size(); ++q) {
            const int i = order[q];
            const int sz = padded_size(pb.ops[i].ncut);
            if(off + sz > limit_doubles || pb.ops[i].ncut > maxcut_smem) break;
            pr.mat_off[i] = off; pr.mat_pad[i] = 1;
            off += sz;
        }
        pr.smem_doubles = off;
        off = round_up(off, 16);
        if(q == order.size()) off = pr.smem_doubles;
        const int gbase = off;
        for(; q < order.size(); ++q) {
            const int i = order[q];
            pr.mat_off[i] = off; pr.mat_pad[i] = 0;
            off += round_up(1 << (2 * pb.ops[i].ncut), 16);
        }
        // the global part starts at arena offset smem_doubles in the kernels' address arithmetic
        if(gbase != pr.smem_doubles) for(size_t r = 0; r < order.size(); ++r) {
            const int i = order[r];
            if(!pr.mat_pad[i]) pr.mat_off[i] -= gbase - pr.smem_doubles;
        }
        pr.arena_doubles = off - (gbase - pr.smem_doubles);
    }
    pr.team_smem_bytes = slk_team_layout(N, pb.n_founders, nops, pr.smem_doubles, tpc_doubles).total;

    // ---- stream ----
    pr.stream.assign((size_t) stream_words, 0u);
    pr.dbg.assign((size_t) nops * 3, 0u);
    for(int i = 0; i < nops; ++i) {
        const slk_peel_op& op = pb.ops[i];
        uint32_t* rec = &pr.stream[(size_t) pr.op_start[i] * 4];
        const int c = op.ncut;
        int nkids = (op.type == SLK_CHILD_PEEL) ? 1 : (op.type == SLK_PARENT_PEEL ? op.nchild : 0);

        bool peel_in_prev = false;
        for(int j = 0; j < op.nprev; ++j) {
            const slk_peel_op& pv = pb.ops[op.prev[j]];
            for(int d = 0; d < pv.ncut; ++d) if(pv.cutset[d] == op.peelnode) peel_in_prev = true;
        }
        rec[0] = (uint32_t) op.type | ((uint32_t) c << 4) | ((uint32_t) op.nprev << 8) | ((uint32_t) nkids << 12) |
                 ((uint32_t) op.peelnode << 16);
        rec[1] = (uint32_t) pr.mat_off[i] | (pr.mat_pad[i] ? SLK_MAT_PAD : 0u);
        rec[2] = qstart[i];
        rec[3] = (peel_in_prev ? 1u : 0u) | ((trait ? (uint32_t) dmask[op.peelnode] : 15u) << 8) | ((uint32_t) i << 16);
        uint32_t* cw = rec + SLK_REC_HEADER;
        for(int d = 0; d < c; ++d) cw[d >> 1] |= (uint32_t) scut[i][d] << (16 * (d & 1));
        uint32_t* pw = cw + ((c + 1) >> 1);
        if(trait) {
            pw[0] = (uint32_t)(static_lmw[i] & 0xffffffffu);
            pw[1] = (uint32_t)(static_lmw[i] >> 32);
            pw += 2;
        }
        uint64_t refpos = 0;
        for(int d = 0; d < c; ++d) refpos |= (uint64_t) find_pos(op, scut[i][d]) << (4 * d);
        pr.dbg[3 * (size_t) i + 0] = (uint32_t) hp.dense_off[i];
        pr.dbg[3 * (size_t) i + 1] = (uint32_t)(refpos & 0xffffffffu);
        pr.dbg[3 * (size_t) i + 2] = (uint32_t)(refpos >> 32);
        for(int j = 0; j < op.nprev; ++j, pw += SLK_REC_PREV) {
            const int x = op.prev[j];
            const std::vector<int>& xs = scut[x];
            // sorted layout: the consumer's peel node, if the consumed function is keyed on it, is its digit 0
            // and the consumer's digit 0 comes next; every other digit is one of the consumer's row digits
            size_t first = 0;
            uint32_t kind;
            const bool has_v = !xs.empty() && xs[0] == op.peelnode;
            if(has_v) first = 1;
            const bool has_s = c > 0 && xs.size() > first && xs[first] == scut[i][0];
            if(has_s) first += 1;
            kind = has_v ? (has_s ? SLK_KEY_VS : SLK_KEY_V) : (has_s ? SLK_KEY_S : SLK_KEY_R);
            // (digit of the consumed matrix's row index) <- (row digit of the consumer's cell), merged into runs
            std::vector<std::pair<int,int> > maps;
            for(size_t r = first; r < xs.size(); ++r) {
                const int pos = spos(i, xs[r]);
                if(pos < 1 || pos == SLK_POS_PEEL) {
                    std::ostringstream e; e << "op " << i << ": previous function " << x << " breaks the sorted-digit layout";
                    err = e.str(); return false;
                }
                maps.push_back(std::make_pair((int)(r - first), pos));
            }
            std::vector<uint32_t> runs;
            for(size_t q = 0; q < maps.size(); ) {
                size_t e = q + 1;
                while(e < maps.size() && maps[e].first == maps[e-1].first + 1 && maps[e].second == maps[e-1].second + 1) ++e;
                const uint32_t nbits = 2 * (uint32_t)(e - q);
                runs.push_back((uint32_t)(2 * maps[q].second) | ((uint32_t)(2 * maps[q].first) << 5) | (nbits << 10));
                q = e;
            }
            if(runs.size() > 5) { err = "gather of a previous function needs more than five runs"; return false; }
            while(runs.size() < 5) runs.push_back(0u);
            size_t nruns = 0;
            for(size_t q = 0; q < 5; ++q) if(runs[q]) nruns = q + 1;
            pw[0] = (uint32_t) pr.mat_off[x] | (pr.mat_pad[x] ? SLK_MAT_PAD : 0u);
            pw[1] = kind | ((uint32_t) nruns << 8) | (runs[4] << 16);
            pw[2] = runs[0] | (runs[1] << 16);
            pw[3] = runs[2] | (runs[3] << 16);
        }
        for(int k = 0; k < nkids; ++k) {
            int kid = (op.type == SLK_CHILD_PEEL) ? op.peelnode : op.children[k];
            int kid_pos = spos(i, kid);
            int mat_pos = spos(i, pb.mother[kid]);
            int pat_pos = spos(i, pb.father[kid]);
            uint32_t male = (pb.sex[kid] == SLK_MALE) ? 1u : 0u;
            pw[k] = (uint32_t) kid | ((uint32_t) kid_pos << 16) | ((uint32_t) mat_pos << 20) |
                    ((uint32_t) pat_pos << 24) | (male << 28);
        }
    }
    // ---- the device image ----
    pr.blob.assign((size_t) lay.blob_bytes, 0);
    memcpy(&pr.blob[lay.off_stream], pr.stream.data(), pr.stream.size() * 4);
    memcpy(&pr.blob[lay.off_op_start], pr.op_start.data(), pr.op_start.size() * 2);
    memcpy(&pr.blob[lay.off_imap], pr.imap.data(), pr.imap.size() * 2);
    memcpy(&pr.blob[lay.off_flevel_items], pr.flevel_items.data(), pr.flevel_items.size() * 4);
    memcpy(&pr.blob[lay.off_flevel_map], pr.flevel_map.data(), pr.flevel_map.size() * 4);
    memcpy(&pr.blob[lay.off_bops], pr.bops.data(), pr.bops.size() * 2);
    memcpy(&pr.blob[lay.off_blevel_start], pr.blevel_start.data(), pr.blevel_start.size() * 2);
    memcpy(&pr.blob[lay.off_glist], kSlkGlist, 16);
    if(trait) memcpy(&pr.blob[lay.off_dprob], pb.disease_prob, (size_t) N * 32);
    lay.stream_words = stream_words;
    lay.trait = trait ? 1 : 0;
    lay.imap_len = (int) pr.imap.size();
    lay.n_flevels = n_flevels;
    lay.n_blevels = n_blevels;
    lay.arena_doubles = pr.arena_doubles;
    lay.smem_doubles = pr.smem_doubles;
    lay.team_threads = pr.team_threads;
    lay.cta_threads = pr.cta_threads;
    lay.prog_smem_bytes = pr.prog_smem_bytes;
    lay.team_smem_bytes = pr.team_smem_bytes;
    lay.table_doubles_per_child = pr.table_doubles_per_child;
    return true;
}

bool build_plan(const slk_problem& pb, HostPlan& hp, std::string& err) {
    std::ostringstream e;
    const int N = pb.n_members, F = pb.n_founders, M = pb.n_markers;
    if(N < 1 || N > 65535 || F < 1 || F > N || M < 2 || pb.n_lod < 1 || pb.n_ops != N) {
        e << "bad dimensions (N=" << N << " F=" << F << " M=" << M << " n_lod=" << pb.n_lod << " n_ops=" << pb.n_ops << ")";
        err = e.str(); return false;
    }
    if(!pb.mother || !pb.father || !pb.sex || !pb.typed || !pb.genotypes || !pb.disease_prob || !pb.marker_prob ||
       !pb.marker_xprob || !pb.theta || !pb.partial_theta || !pb.elimination || !pb.ops) {
        err = "null array in slk_problem"; return false;
    }
    for(int i = 0; i < N; ++i) {
        bool founder = pb.mother[i] < 0 && pb.father[i] < 0;
        if(founder != (i < F)) { e << "person " << i << ": founders must come first"; err = e.str(); return false; }
        if(!founder && (pb.mother[i] < 0 || pb.father[i] < 0 || pb.mother[i] >= N || pb.father[i] >= N)) {
            e << "person " << i << ": bad parent ids"; err = e.str(); return false;
        }
    }

    hp.N = N; hp.F = F; hp.M = M; hp.nlod = pb.n_lod; hp.sex_linked = pb.sex_linked ? 1 : 0; hp.nops = pb.n_ops;
    hp.mother.resize(N); hp.father.resize(N); hp.male.resize(N);
    for(int i = 0; i < N; ++i) {
        hp.mother[i] = (int16_t) pb.mother[i];
        hp.father[i] = (int16_t) pb.father[i];
        hp.male[i] = pb.sex[i] == SLK_MALE;
    }
    hp.disease_prob.assign(pb.disease_prob, pb.disease_prob + 4 * N);
    const double* pp = pb.person_prior ? pb.person_prior : pb.disease_prob;
    hp.person_prior.assign(pp, pp + 4 * N);

    // ---- peel sequence sanity ----
    std::vector<int> seen(N, 0), used(N, 0);
    hp.dense_off.resize(N);
    hp.sum_cells = hp.sum_presum = 0;
    hp.max_cutset = 0;
    hp.flops_ls = hp.flops_lod = 0.0;
    int n_child_tables = 0;
    for(int i = 0; i < N; ++i) {
        const slk_peel_op& op = pb.ops[i];
        if(op.peelnode < 0 || op.peelnode >= N || seen[op.peelnode]) { e << "op " << i << ": bad peel node"; err = e.str(); return false; }
        seen[op.peelnode] = 1;
        if(op.ncut < 0 || op.ncut > SLK_MAX_CUTSET) { e << "op " << i << ": cutset of " << op.ncut << " exceeds SLK_MAX_CUTSET"; err = e.str(); return false; }
        if(op.nprev < 0 || op.nprev > SLK_MAX_PREV || op.nchild < 0 || op.nchild > SLK_MAX_CHILDREN) { e << "op " << i << ": too many previous functions / children"; err = e.str(); return false; }
        if(op.type < SLK_CHILD_PEEL || op.type > SLK_LAST_PEEL) { e << "op " << i << ": bad type"; err = e.str(); return false; }
        for(int d = 0; d < op.ncut; ++d) {
            if(op.cutset[d] < 0 || op.cutset[d] >= N || op.cutset[d] == op.peelnode) { e << "op " << i << ": bad cutset"; err = e.str(); return false; }
        }
        for(int j = 0; j < op.nprev; ++j) {
            int q = op.prev[j];
            if(q < 0 || q >= i || used[q]) { e << "op " << i << ": previous function " << q << " invalid or consumed twice"; err = e.str(); return false; }
            used[q] = 1;
            for(int d = 0; d < pb.ops[q].ncut; ++d) {
                if(find_pos(op, pb.ops[q].cutset[d]) < 0) { e << "op " << i << ": previous function " << q << " is keyed on a person outside cutset+peelnode"; err = e.str(); return false; }
            }
        }
        if(op.type == SLK_CHILD_PEEL) {
            int kid = op.peelnode;
            if(kid < F || find_pos(op, pb.mother[kid]) < 0 || find_pos(op, pb.father[kid]) < 0) { e << "op " << i << ": child peel without both parents in the cutset"; err = e.str(); return false; }
        }
        if(op.type == SLK_PARENT_PEEL) {
            // the trait R-function walks the cutset (trait_rfunction.cc:97-104), the sampler the
            // children list (sampler_rfunction.cc:263-265): they must be the same sequence
            int k = 0;
            for(int d = 0; d < op.ncut; ++d) {
                int q = op.cutset[d];
                bool is_child = q >= F && (pb.mother[q] == op.peelnode || pb.father[q] == op.peelnode);
                if(is_child) {
                    if(k >= op.nchild || op.children[k] != q) { e << "op " << i << ": children list disagrees with the cutset"; err = e.str(); return false; }
                    int other = (pb.mother[q] == op.peelnode) ? pb.father[q] : pb.mother[q];
                    if(find_pos(op, other) < 0 || find_pos(op, other) == SLK_POS_PEEL) { e << "op " << i << ": other parent of child " << q << " not in the cutset"; err = e.str(); return false; }
                    ++k;
                }
            }
            if(k != op.nchild) { e << "op " << i << ": children list disagrees with the cutset"; err = e.str(); return false; }
        }
        hp.dense_off[i] = (int) hp.sum_cells;
        hp.sum_cells += 1L << (2 * op.ncut);
        hp.sum_presum += 1L << (2 * op.ncut + 2);
        hp.max_cutset = std::max(hp.max_cutset, op.ncut);

        // algorithmic work without validity pruning
        int t = (op.type == SLK_CHILD_PEEL) ? 1 : (op.type == SLK_PARENT_PEEL ? op.nchild : 0);
        hp.flops_ls += (double)(1L << (2 * op.ncut + 2)) * (1 + op.nprev + t);
        double w = (op.type == SLK_CHILD_PEEL) ? 4.0 * (2 + op.nprev + 4)
                 : (op.type == SLK_PARENT_PEEL) ? 4.0 * (1 + op.nprev + op.nchild * (4 * 5 + 1))
                 : 4.0 * (1 + op.nprev);
        hp.flops_lod += (double)(1L << (2 * op.ncut)) * w;
        if(op.type == SLK_CHILD_PEEL) n_child_tables += 1;
        if(op.type == SLK_PARENT_PEEL) n_child_tables += op.nchild;
    }
    if(pb.ops[N-1].ncut != 0) { err = "last op must have an empty cutset"; return false; }
    hp.last_op = N - 1;
    hp.flops_ls += 64.0 * 6.0 * n_child_tables + 8.0 * N + 6.0 * (N - F);
    hp.flops_lod += 2.0 * (N - F);

    // ---- per (locus, person) code: legal mask | prior class ----
    hp.gcode.resize((size_t) M * N);
    hp.fprior.resize((size_t) M * 8);
    for(int l = 0; l < M; ++l) {
        for(int i = 0; i < N; ++i) {
            bool founder = pb.prior_as_founder ? (pb.prior_as_founder[i] != 0) : true;
            bool xmale = pb.sex_linked && pb.sex[i] == SLK_MALE;
            int g = pb.genotypes[(size_t) i * M + l];
            int cls = prior_class(founder, pb.typed[i] != 0, g, xmale);
            if(pb.disease_prior_locus_plus1 == l + 1) cls = SLK_PRIOR_PERSON;
            uint8_t lm = legal_mask_from_elimination(pb.elimination[(size_t) l * N + i]);
            hp.gcode[(size_t) l * N + i] = (uint8_t)(lm | (cls << 4));
        }
        // class FOUNDER: autosomal / female map prior, or (typed X male, untyped marker) the
        // X-male prior with hetero kept (it is 0 in the table anyway, genetic_map.h:78-86)
        founder_prior(pb.marker_prob + 4 * l, false, &hp.fprior[(size_t) l * 8]);
        founder_prior(pb.marker_xprob + 4 * l, true, &hp.fprior[(size_t) l * 8 + 4]);
    }

    hp.theta.assign(pb.theta, pb.theta + (M - 1));
    hp.partial.assign(pb.partial_theta, pb.partial_theta + (M - 1));
    hp.log_theta.resize(M - 1); hp.log_1mtheta.resize(M - 1);
    for(int l = 0; l < M - 1; ++l) {
        if(!(pb.theta[l] > 0.0 && pb.theta[l] < 1.0)) { e << "theta[" << l << "] = " << pb.theta[l] << " outside (0,1)"; err = e.str(); return false; }
        hp.log_theta[l] = log(pb.theta[l]);                   // genetic_map.cc:132-138
        hp.log_1mtheta[l] = log(1.0 - pb.theta[l]);
    }
    hp.marker_transmission = pb.sex_linked ? log(0.5) * (N - F) : log(0.5) * (2 * (N - F));   // descent_graph.cc:22,35

    // ---- M-sampler tables ----
    hp.ms_available = pb.minor_freq != 0;
    hp.ms_seq.clear(); hp.ms_typed.clear(); hp.ms_ordering.clear();
    {
        // meiosis_sampler.cc:41-72: founders, then repeated passes picking people whose parents are placed
        std::vector<char> visited(N, 0);
        for(int i = 0; i < F; ++i) visited[i] = 1;
        int total = N - F;
        while(total > 0) {
            int placed = 0;
            for(int i = F; i < N; ++i) {
                if(visited[i]) continue;
                if(visited[pb.mother[i]] && visited[pb.father[i]]) { hp.ms_seq.push_back((uint16_t) i); visited[i] = 1; --total; ++placed; }
            }
            if(!placed) { err = "pedigree has a person who is their own ancestor"; return false; }
        }
        for(int i = 0; i < N; ++i) if(pb.typed[i]) hp.ms_typed.push_back((uint16_t) i);
        const int nt = (int) hp.ms_typed.size();
        hp.ms_typed_auto.assign((size_t) std::max(nt, 1), 0);
        for(int k = 0; k < nt; ++k) hp.ms_typed_auto[k] = (pb.sex_linked && pb.sex[hp.ms_typed[k]] == SLK_MALE) ? 1 : 0;
        hp.ms_obsT.resize((size_t) std::max(nt, 1) * M);
        for(int k = 0; k < nt; ++k)
            for(int l = 0; l < M; ++l) hp.ms_obsT[(size_t) k * M + l] = pb.genotypes[(size_t) hp.ms_typed[k] * M + l];
        hp.ms_minor.assign(M, 0.0);
        if(pb.minor_freq) hp.ms_minor.assign(pb.minor_freq, pb.minor_freq + M);
        hp.ms_lnmajor.resize(M); hp.ms_lnminor.resize(M);
        for(int l = 0; l < M; ++l) {
            const double mn = hp.ms_minor[l], mj = 1.0 - mn;                    // genetic_map.h:42-47
            hp.ms_lnmajor[l] = mj > 0.0 ? log(mj) : -1e300;
            hp.ms_lnminor[l] = mn > 0.0 ? log(mn) : -1e300;
        }
        // slot masks for the incremental step kernel: slot 2k + s of typed person k is "below" person P if its
        // lineage can pass through P, i.e. parent_s(typed[k]) is P or a descendant of P
        {
            hp.ms_W = std::max(1, (2 * nt + 31) / 32);
            hp.ms_typed_index.assign(N, -1);
            for(int k = 0; k < nt; ++k) hp.ms_typed_index[hp.ms_typed[k]] = (int16_t) k;
            // anc[i] = set of ancestors-or-self of i among non-founders, as a bit matrix (N-F columns)
            const int NF = N - F, AW = (NF + 31) / 32;
            std::vector<uint32_t> anc((size_t) N * AW, 0u);
            for(size_t q = 0; q < hp.ms_seq.size(); ++q) {
                const int i = hp.ms_seq[q];
                uint32_t* a = &anc[(size_t) i * AW];
                a[(i - F) >> 5] |= 1u << ((i - F) & 31);
                const int par[2] = { pb.mother[i], pb.father[i] };
                for(int s = 0; s < 2; ++s) if(par[s] >= F) for(int w = 0; w < AW; ++w) a[w] |= anc[(size_t) par[s] * AW + w];
            }
            hp.ms_desc_mask.assign((size_t) std::max(NF, 1) * hp.ms_W, 0u);
            for(int k = 0; k < nt; ++k) {
                const int i = hp.ms_typed[k];
                if(i < F) continue;
                const int par[2] = { pb.mother[i], pb.father[i] };
                for(int s = 0; s < 2; ++s) {
                    if(par[s] < F) continue;
                    const uint32_t* a = &anc[(size_t) par[s] * AW];
                    for(int P = 0; P < NF; ++P)
                        if(a[P >> 5] & (1u << (P & 31))) hp.ms_desc_mask[(size_t) P * hp.ms_W + ((2 * k + s) >> 5)] |= 1u << ((2 * k + s) & 31);
                }
            }
        }
        // markov_chain.cc:68-80 with Person::safe_to_ignore_meiosis (person.cc:208-222)
        std::vector<int> nchild(N, 0);
        for(int i = F; i < N; ++i) { nchild[pb.mother[i]]++; nchild[pb.father[i]]++; }
        for(int i = 0; i < 2 * (N - F); ++i) {
            const int person = F + i / 2, par = i % 2;
            const int parent = par == 0 ? pb.mother[person] : pb.father[person];
            bool ignore;
            if(parent >= F) ignore = pb.sex_linked ? (par == 1) : false;
            else ignore = nchild[parent] == 1;
            if(!ignore) hp.ms_ordering.push_back(i);
        }
    }

    hp.ls = HostProgram(); hp.lod = HostProgram();
    hp.ls.arena_doubles = hp.lod.arena_doubles = 0;
    if(!emit_program(pb, hp, false, hp.ls, err)) return false;
    if(!emit_program(pb, hp, true, hp.lod, err)) return false;
    return true;
}

}  // namespace slk
