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

static void emit_program(const slk_problem& pb, const HostPlan& hp, bool trait, HostProgram& pr) {
    const int nops = pb.n_ops;

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

    // ---- backward levels: an op is sampled after the ops that peel its cutset members ----
    std::vector<int> peeled_by(pb.n_members, -1);
    for(int i = 0; i < nops; ++i) peeled_by[pb.ops[i].peelnode] = i;
    std::vector<int> blevel(nops, 0);
    int n_blevels = 0;
    for(int i = nops - 1; i >= 0; --i) {
        int lv = 0;
        for(int d = 0; d < pb.ops[i].ncut; ++d) lv = std::max(lv, blevel[peeled_by[pb.ops[i].cutset[d]]] + 1);
        blevel[i] = lv;
        n_blevels = std::max(n_blevels, lv + 1);
    }

    // ---- static validity: cells per op ----
    std::vector<long> ncells(nops), nrows(nops);
    std::vector<uint64_t> static_lmw(nops, 0);
    std::vector<uint8_t> dmask(pb.n_members, 0);
    for(int i = 0; i < pb.n_members; ++i)
        for(int g = 0; g < 4; ++g) if(pb.disease_prob[i*4 + g] != 0.0) dmask[i] |= (uint8_t)(1 << g);
    pr.valid_cells = 0;
    for(int i = 0; i < nops; ++i) {
        const slk_peel_op& op = pb.ops[i];
        long n = 1;
        for(int d = 0; d < op.ncut; ++d) {
            uint8_t m = trait ? dmask[op.cutset[d]] : 15;
            static_lmw[i] |= (uint64_t) m << (4 * d);
            n *= __builtin_popcount(m);
        }
        ncells[i] = n;
        // rows = valid assignments of digits 1..c-1 (digit 0 is the register tile's slot axis)
        nrows[i] = (op.ncut > 0) ? n / __builtin_popcount((unsigned)(static_lmw[i] & 15u)) : 1;
        pr.valid_cells += n;
    }

    // ---- arena ----
    pr.mat_off.assign(nops, 0);
    std::vector<int> level_order;
    for(int lv = 0; lv < n_flevels; ++lv) for(int i = 0; i < nops; ++i) if(flevel[i] == lv) level_order.push_back(i);

    if(trait) {
        // matrices die after the level of their single consumer
        Liveness lv;
        std::vector<std::vector<int> > dies_after(n_flevels);
        for(int i = 0; i < nops; ++i) if(consumer[i] >= 0) dies_after[flevel[consumer[i]]].push_back(i);
        size_t k = 0;
        for(int l = 0; l < n_flevels; ++l) {
            for(; k < level_order.size() && flevel[level_order[k]] == l; ++k) {
                int i = level_order[k];
                pr.mat_off[i] = lv.alloc(1 << (2 * pb.ops[i].ncut));
            }
            for(size_t q = 0; q < dies_after[l].size(); ++q) {
                int i = dies_after[l][q];
                lv.release(pr.mat_off[i], 1 << (2 * pb.ops[i].ncut));
            }
        }
        int hw = 0;
        for(int i = 0; i < nops; ++i) hw = std::max(hw, pr.mat_off[i] + (1 << (2 * pb.ops[i].ncut)));
        pr.arena_doubles = hw;
    }
    else {
        // everything stays live until the backward pass; large matrices go last so that, if the
        // arena outgrows shared memory, it is the few big ones that land in the global slab
        std::vector<int> order(nops);
        for(int i = 0; i < nops; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(),
                         [&](int a, int b) { return pb.ops[a].ncut < pb.ops[b].ncut; });
        int off = 0;
        for(int q = 0; q < nops; ++q) {
            pr.mat_off[order[q]] = off;
            off += 1 << (2 * pb.ops[order[q]].ncut);
        }
        pr.arena_doubles = off;
    }

    // ---- forward schedule: ops per level, quads (4 cells per thread) per op ----
    pr.lops.clear(); pr.lpf.clear(); pr.flevel_quads.clear();
    pr.flevel_start.assign(1, 0);
    for(int l = 0; l < n_flevels; ++l) {
        // big ops first inside a level so that the threads of a warp mostly share an op
        std::vector<int> ops_here;
        for(int i = 0; i < nops; ++i) if(flevel[i] == l) ops_here.push_back(i);
        // ... and ops of the same shape next to each other so that the threads of a warp that
        // do hold different ops still follow the same control flow
        auto shape = [&](int i) {
            const slk_peel_op& o = pb.ops[i];
            return std::make_tuple(-nrows[i], o.type, o.nprev, o.nchild, o.ncut);
        };
        std::stable_sort(ops_here.begin(), ops_here.end(), [&](int a, int b) { return shape(a) < shape(b); });
        uint32_t quads = 0;
        for(size_t q = 0; q < ops_here.size(); ++q) {
            pr.lops.push_back((uint16_t) ops_here[q]);
            pr.lpf.push_back(quads);
            quads += (uint32_t) nrows[ops_here[q]];
        }
        pr.flevel_quads.push_back(quads);
        pr.flevel_start.push_back((uint16_t) pr.lops.size());
    }

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

    // ---- stream ----
    pr.stream.clear();
    pr.op_start.assign(nops, 0);
    for(int i = 0; i < nops; ++i) {
        const slk_peel_op& op = pb.ops[i];
        pr.op_start[i] = (uint16_t) pr.stream.size();
        int nkids = 0;
        if(op.type == SLK_CHILD_PEEL) nkids = 1;
        else if(op.type == SLK_PARENT_PEEL) nkids = op.nchild;

        bool peel_in_prev = false;
        for(int j = 0; j < op.nprev; ++j) {
            const slk_peel_op& pv = pb.ops[op.prev[j]];
            for(int d = 0; d < pv.ncut; ++d) if(pv.cutset[d] == op.peelnode) peel_in_prev = true;
        }

        pr.stream.push_back((uint32_t) op.type | ((uint32_t) op.ncut << 4) | ((uint32_t) op.nprev << 8) |
                            ((uint32_t) nkids << 12) | ((uint32_t) op.peelnode << 16));
        pr.stream.push_back((uint32_t) pr.mat_off[i]);
        pr.stream.push_back((uint32_t) hp.dense_off[i]);
        pr.stream.push_back(peel_in_prev ? 1u : 0u);
        pr.stream.push_back((uint32_t)(static_lmw[i] & 0xffffffffu));
        pr.stream.push_back((uint32_t)(static_lmw[i] >> 32));
        pr.stream.push_back(trait ? (uint32_t) dmask[op.peelnode] : 15u);
        for(int d = 0; d < op.ncut; d += 2) {
            uint32_t w = (uint32_t) op.cutset[d];
            if(d + 1 < op.ncut) w |= (uint32_t) op.cutset[d + 1] << 16;
            pr.stream.push_back(w);
        }
        for(int j = 0; j < op.nprev; ++j) {
            const slk_peel_op& pv = pb.ops[op.prev[j]];
            uint32_t peel_shift = SLK_NO_SHIFT, d0_shift = SLK_NO_SHIFT;
            // (dst digit in the consumed matrix) <- (src digit of the consumer's cutset);
            // digit 0 of the consumer is the register tile's slot axis and is kept out of the runs
            std::vector<std::pair<int,int> > maps;
            for(int d = 0; d < pv.ncut; ++d) {
                int pos = find_pos(op, pv.cutset[d]);
                if(pos == SLK_POS_PEEL) peel_shift = 2 * d;
                else if(pos == 0) d0_shift = 2 * d;
                else maps.push_back(std::make_pair(d, pos));
            }
            std::vector<uint32_t> runs;
            for(size_t q = 0; q < maps.size(); ) {
                size_t e = q + 1;
                while(e < maps.size() && maps[e].first == maps[e-1].first + 1 && maps[e].second == maps[e-1].second + 1) ++e;
                uint32_t nbits = 2 * (uint32_t)(e - q);
                runs.push_back((uint32_t)(2 * maps[q].second) | ((uint32_t)(2 * maps[q].first) << 5) | (nbits << 10));
                q = e;
            }
            pr.stream.push_back((uint32_t) pr.mat_off[op.prev[j]]);
            pr.stream.push_back(peel_shift | ((uint32_t) runs.size() << 8) | (d0_shift << 16));
            for(size_t q = 0; q < runs.size(); q += 2) {
                uint32_t w = runs[q];
                if(q + 1 < runs.size()) w |= runs[q + 1] << 16;
                pr.stream.push_back(w);
            }
        }
        for(int k = 0; k < nkids; ++k) {
            int kid = (op.type == SLK_CHILD_PEEL) ? op.peelnode : op.children[k];
            int kid_pos = find_pos(op, kid);
            int mat_pos = find_pos(op, pb.mother[kid]);
            int pat_pos = find_pos(op, pb.father[kid]);
            uint32_t male = (pb.sex[kid] == SLK_MALE) ? 1u : 0u;
            pr.stream.push_back((uint32_t) kid | ((uint32_t) kid_pos << 16) | ((uint32_t) mat_pos << 20) |
                                ((uint32_t) pat_pos << 24) | (male << 28));
        }
    }

    // ---- geometry ----
    const int N = pb.n_members;
    pr.prog_smem_bytes = round_up((int)(pr.stream.size() * 4 + pr.lpf.size() * 4 + pr.flevel_quads.size() * 4 +
                                        pr.op_start.size() * 2 + pr.lops.size() * 2 +
                                        pr.flevel_start.size() * 2 + pr.bops.size() * 2 +
                                        pr.blevel_start.size() * 2 + 64), 16);
    const int tpc_doubles = trait ? 4 : 16;
    pr.table_doubles_per_child = tpc_doubles;
    const int table_bytes = slk_team_layout(N, pb.n_founders, nops, 0, tpc_doubles).total;
    const long work = trait ? pr.valid_cells : hp.sum_cells;
    int team = 32;
    if(work > 1024)  team = 64;
    if(work > 4096)  team = 128;
    if(work > 8192)  team = 256;
    // One unit is latency bound (a chain of ~50 dependent levels), so the SM needs several units in
    // flight.  When the whole arena of a team would crowd the others out of shared memory, only the
    // small matrices (cutset <= 4, the ones hit by the many small latency-bound ops) stay in shared
    // memory and the few large ones go to the L2-resident global slab (measured on B200 with the
    // 200-member pedigree: 8.4 ms per sweep against 11.5 ms with everything in shared memory).
    int small_doubles = 0;
    for(int i = 0; i < nops; ++i) if(pb.ops[i].ncut <= 4) small_doubles = std::max(small_doubles, pr.mat_off[i] + (1 << (2 * pb.ops[i].ncut)));
    const bool hybrid = (size_t) pr.arena_doubles * 8 + table_bytes > 100 * 1024;
    team = env_int(trait ? "SLK_LOD_TEAM" : "SLK_LS_TEAM", team);
    if(team != 32 && team != 64 && team != 128 && team != 256 && team != 512) team = 128;
    pr.team_threads = team;
    pr.cta_threads = std::max(team, env_int(trait ? "SLK_LOD_CTA_THREADS" : "SLK_LS_CTA_THREADS", env_int("SLK_CTA_THREADS", hybrid ? 768 : 128)));

    const int teams_per_cta = pr.cta_threads / team;
    const int budget = kSmemPerBlockMax - pr.prog_smem_bytes;           // one CTA per SM worst case
    int per_team_budget = budget / teams_per_cta - table_bytes;
    if(per_team_budget < 0) per_team_budget = 0;
    int smem_doubles = std::min(pr.arena_doubles, per_team_budget / 8);
    if(hybrid && !trait) smem_doubles = std::min(smem_doubles, small_doubles);
    smem_doubles = env_int(trait ? "SLK_LOD_SMEM_DOUBLES" : "SLK_LS_SMEM_DOUBLES", smem_doubles);
    smem_doubles = std::min(smem_doubles, std::min(pr.arena_doubles, per_team_budget / 8));
    if(smem_doubles < pr.arena_doubles) {
        // only whole matrices live in shared memory: cut at a matrix boundary
        int cut = 0;
        std::vector<std::pair<int,int> > spans;
        for(int i = 0; i < nops; ++i) spans.push_back(std::make_pair(pr.mat_off[i], 1 << (2 * pb.ops[i].ncut)));
        std::sort(spans.begin(), spans.end());
        for(size_t q = 0; q < spans.size(); ++q) {
            // (the trait arena reuses space, so spans may overlap; a boundary is any span start
            //  not inside an earlier span that crosses the cut)
            if(spans[q].first + spans[q].second <= smem_doubles) cut = std::max(cut, spans[q].first + spans[q].second);
        }
        // make sure no span straddles the cut
        bool ok = false;
        while(!ok) {
            ok = true;
            for(size_t q = 0; q < spans.size(); ++q) {
                if(spans[q].first < cut && spans[q].first + spans[q].second > cut) { cut = spans[q].first; ok = false; }
            }
        }
        smem_doubles = cut;
    }
    pr.smem_doubles = smem_doubles;
    pr.team_smem_bytes = slk_team_layout(N, pb.n_founders, nops, smem_doubles, tpc_doubles).total;
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
    emit_program(pb, hp, false, hp.ls);
    emit_program(pb, hp, true, hp.lod);
    if(hp.ls.stream.size() > 65535 || hp.lod.stream.size() > 65535) {
        err = "peel program too large for 16-bit offsets"; return false;
    }
    return true;
}

}  // namespace slk
