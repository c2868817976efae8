// slk_emu.cc -- sequential CPU emulation of the peel kernels' per-thread code (slk_peel.h) over the flattened
// plan (slk_plan.cc).  TEST INFRASTRUCTURE ONLY: it lets the CPU test-suite check the plan flattening, the
// sorted-digit matrix layout and the tile / backward index arithmetic bit for bit against the oracle without a
// GPU.  It is built into tests/emu/libslk_emu.so by tests/emu/build.py, never into the product library, and
// nothing under swiftlink_b200/ loads it.  The orchestration below mirrors slk_lsampler_kernel /
// slk_lodscore_kernel (slk_kernels.cuh) with the team's threads run one after another.
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "swiftlink_b200.h"
#include "slk_plan.h"
#include "slk_peel.h"
#include "slk_philox.cuh"

namespace {

struct Emu {
    slk::HostPlan hp;
    std::string err;
};

struct TeamMem {
    std::vector<unsigned char> smem;     // the team's shared-memory slab, laid out by slk_team_layout
    std::vector<double> garena;
    SlkTeam tm;
};

SlkProgView view_of(const slk::HostProgram& pr, const double* dprob_global) {
    const unsigned char* b = pr.blob.data();
    const SlkProgram& p = pr.layout;
    SlkProgView g;
    g.stream = (const uint32_t*)(b + p.off_stream);
    g.op_start = (const uint16_t*)(b + p.off_op_start);
    g.imap = (const uint16_t*)(b + p.off_imap);
    g.flevel_items = (const uint32_t*)(b + p.off_flevel_items);
    g.flevel_map = (const uint32_t*)(b + p.off_flevel_map);
    g.bops = (const uint16_t*)(b + p.off_bops);
    g.blevel_start = (const uint16_t*)(b + p.off_blevel_start);
    g.glist = b + p.off_glist;
    g.dprob = p.off_dprob >= 0 ? (const double*)(b + p.off_dprob) : dprob_global;
    return g;
}

void carve(const slk::HostPlan& hp, const slk::HostProgram& pr, TeamMem& t) {
    SlkTeamLayout lay = slk_team_layout(hp.N, hp.F, hp.nops, pr.smem_doubles, pr.table_doubles_per_child);
    t.smem.assign((size_t) lay.total + 16, 0);
    // poison what the kernels do not zero themselves: reading an unstaged value must show up
    t.garena.assign((size_t)(pr.arena_doubles - pr.smem_doubles) + 16, 0.0);
    unsigned char* base = t.smem.data();
    while(((uintptr_t) base) & 15) ++base;
    t.tm.arena = (double*)(base + lay.arena);
    t.tm.smem_doubles = pr.smem_doubles;
    t.tm.garena = t.garena.data();
    t.tm.tables = (double*)(base + lay.tables);
    t.tm.scal = (double*)(base + lay.scal);
    t.tm.oploc = (SlkOpLoc*)(base + lay.oploc);
    t.tm.ru = (double*)(base + lay.ru);
    t.tm.gc = base + lay.bytes;
    t.tm.dgl = t.tm.gc + lay.nbytes;
    t.tm.dgr = t.tm.dgl + lay.nbytes;
    t.tm.pmk = t.tm.dgr + lay.nbytes;
    t.tm.red = (int*)(base + lay.red);
}

SlkEnv env_of(const slk::HostPlan& hp, double* dump_mat, double* dump_pre) {
    SlkEnv E;
    E.N = hp.N; E.F = hp.F; E.nops = hp.nops; E.sex_linked = hp.sex_linked;
    E.male = hp.male.data();
    E.dump_mat = dump_mat; E.dump_pre = dump_pre;
    E.dbg = hp.ls.dbg.data();
    return E;
}

uint8_t dg_byte(const int32_t* dg, int N, int locus, int person) {
    const int32_t* p = dg + ((size_t) locus * N + person) * 2;
    return (uint8_t)((p[0] & 1) | ((p[1] & 1) << 1));
}

template<bool TRAIT>
void forward(const slk::HostProgram& pr, const SlkEnv& E, const SlkProgView& pg, const SlkTeam& tm) {
    for(int lv = 0; lv < pr.layout.n_flevels; ++lv) {
        const uint32_t items = pg.flevel_items[lv];
        const uint32_t mw = pg.flevel_map[lv];
        const uint16_t* map = pg.imap + (mw & ~SLK_LEVEL_FINE);
        for(uint32_t q = 0; q < items; ++q) slk_forward_item<TRAIT, true>(E, pg, tm, map, (mw & SLK_LEVEL_FINE) != 0u, q);
    }
}

}  // namespace

extern "C" {

void* emu_create(const slk_problem* pb) {
    Emu* e = new Emu();
    if(!slk::build_plan(*pb, e->hp, e->err)) {
        fprintf(stderr, "emu_create: %s\n", e->err.c_str());
        delete e;
        return 0;
    }
    return e;
}

void emu_destroy(void* h) { delete (Emu*) h; }

int emu_stat(void* h, int which) {
    Emu* e = (Emu*) h;
    switch(which) {
        case 0: return e->hp.ls.smem_doubles;
        case 1: return e->hp.ls.arena_doubles;
        case 2: return e->hp.lod.smem_doubles;
        case 3: return e->hp.lod.arena_doubles;
        case 4: return e->hp.ls.prog_smem_bytes;
        case 5: return e->hp.ls.team_smem_bytes;
        case 6: return e->hp.lod.prog_smem_bytes;
        case 7: return e->hp.lod.team_smem_bytes;
        case 8: return e->hp.ls.team_threads;
        case 9: return e->hp.ls.cta_threads;
        case 10: return e->hp.lod.team_threads;
        case 11: return e->hp.lod.cta_threads;
    }
    return -1;
}

// LocusSampler::step at one locus (slk_lsampler_kernel's unit); dg int32[M][N][2] is updated in place unless
// forward_only.  Returns the likelihood of the locus.
double emu_ls_step(void* h, int32_t* dg, int locus, int ign_left, int ign_right, int forward_only,
                   uint64_t seed, uint32_t chain, uint64_t iteration,
                   double* matrices, double* presums, int32_t* pmk_out, double* dist4) {
    Emu* e = (Emu*) h;
    const slk::HostPlan& hp = e->hp;
    const slk::HostProgram& pr = hp.ls;
    const int N = hp.N, F = hp.F, M = hp.M;
    TeamMem t;
    carve(hp, pr, t);
    const SlkTeam& tm = t.tm;
    const SlkProgView pg = view_of(pr, hp.person_prior.data());
    const SlkEnv E = env_of(hp, matrices, presums);
    const bool has_left = locus != 0, has_right = locus != M - 1;

    for(int i = 0; i < N; ++i) {
        tm.gc[i] = hp.gcode[(size_t) locus * N + i];
        tm.dgl[i] = has_left ? dg_byte(dg, N, locus - 1, i) : 0;
        tm.dgr[i] = has_right ? dg_byte(dg, N, locus + 1, i) : 0;
    }
    for(int i = 0; i < 28; ++i)
        tm.scal[16 + i] = (i < 20) ? kSlkClassPrior[i >> 2][i & 3] : hp.fprior[(size_t) locus * 8 + (i - 20)];
    for(int op = 0; op < hp.nops; ++op) tm.ru[op] = slk_uniform(seed, chain, iteration, (uint32_t) locus, (uint32_t) op);
    {
        double th2 = 1.0, ath2 = 1.0, th = 1.0, ath = 1.0;
        if(has_left && !ign_left)   { th2 = hp.theta[locus - 1]; ath2 = 1.0 - th2; }
        if(has_right && !ign_right) { th = hp.theta[locus]; ath = 1.0 - th; }
        tm.scal[0] = th; tm.scal[1] = ath; tm.scal[2] = th2; tm.scal[3] = ath2;
    }
    for(int q = 0; q < 2 * (N - F); ++q) slk_stage_transmission(E, tm, q, has_left, has_right);
    for(int op = 0; op < hp.nops; ++op) slk_stage_oploc<false>(pg, tm, op);

    forward<false>(pr, E, pg, tm);
    const double result = *slk_mat_ptr(tm, pg.stream[4u * pg.op_start[hp.last_op] + 1u]);
    if(forward_only || result == 0.0) return result;

    for(int lv = 0; lv < pr.layout.n_blevels; ++lv) {
        const int b = pg.blevel_start[lv], en = pg.blevel_start[lv + 1];
        // every op of a level reads the genotypes sampled by earlier levels only: two passes, as the team's
        // barrier separates them on the device
        std::vector<uint8_t> chosen((size_t)(en - b));
        for(int q = b; q < en; ++q) {
            const int op = pg.bops[q];
            double dd[4];
            for(uint32_t g = 0; g < 4; ++g) dd[g] = slk_backward_term(E, pg, tm, (uint32_t) op, g);
            if(dist4) for(int k = 0; k < 4; ++k) dist4[4 * op + k] = dd[k];
            chosen[(size_t)(q - b)] = (uint8_t) slk_sample4(dd, tm.ru[op]);
        }
        for(int q = b; q < en; ++q) {
            const int op = pg.bops[q];
            tm.pmk[pg.stream[4u * pg.op_start[op]] >> 16] = chosen[(size_t)(q - b)];
        }
    }
    const bool use_left = has_left && !ign_left, use_right = has_right && !ign_right;
    const double th_left = use_left ? hp.theta[locus - 1] : 0.0, th_right = use_right ? hp.theta[locus] : 0.0;
    for(int i = F; i < N; ++i) {
        const uint32_t out = slk_sample_indicators(E, tm, i, hp.mother[i], hp.father[i], use_left, use_right, th_left, th_right,
            [&](int parent) { return slk_uniform(seed, chain, iteration, (uint32_t) locus, (uint32_t)(hp.nops + 2 * i + parent)); });
        dg[((size_t) locus * N + i) * 2 + 0] = (int32_t)(out & 1u);
        dg[((size_t) locus * N + i) * 2 + 1] = (int32_t)((out >> 1) & 1u);
    }
    if(pmk_out) for(int i = 0; i < N; ++i) pmk_out[i] = tm.pmk[i];
    return result;
}

// Peeler::process at one (interval, position) (slk_lodscore_kernel's unit); dg == NULL: no descent graph (P(T)).
// Returns the trait likelihood; *prob = ln L - recomb - transmission.
double emu_lod_position(void* h, const int32_t* dg, int interval, int k, double* prob, double* matrices) {
    Emu* e = (Emu*) h;
    const slk::HostPlan& hp = e->hp;
    const slk::HostProgram& pr = hp.lod;
    const int N = hp.N, F = hp.F;
    const bool sex_linked = hp.sex_linked != 0;
    const double trait_prob = sex_linked ? 0.5 : 0.25;
    TeamMem t;
    carve(hp, pr, t);
    const SlkTeam& tm = t.tm;
    const SlkProgView pg = view_of(pr, hp.disease_prob.data());
    const SlkEnv E = env_of(hp, matrices, 0);
    for(int op = 0; op < hp.nops; ++op) slk_stage_oploc<true>(pg, tm, op);
    int ncross = 0;
    if(dg) {
        for(int i = 0; i < N; ++i) { tm.dgl[i] = dg_byte(dg, N, interval, i); tm.dgr[i] = dg_byte(dg, N, interval + 1, i); }
        const double th = hp.partial[interval] * (double)(k + 1);
        const double th2 = hp.partial[interval] * (double)(hp.nlod + 1 - (k + 1));
        tm.scal[0] = th; tm.scal[1] = 1.0 - th; tm.scal[2] = th2; tm.scal[3] = 1.0 - th2;
        for(int q = 0; q < 4 * (N - F); ++q) slk_stage_trait_weight(E, tm, q, trait_prob);
        for(int i = F; i < N; ++i) {
            const uint32_t x = tm.dgl[i] ^ tm.dgr[i];
            ncross += (int)(x & 1u) + (sex_linked ? 0 : (int)((x >> 1) & 1u));
        }
    }
    else for(int q = 0; q < 4 * (N - F); ++q) tm.tables[q] = trait_prob;
    forward<true>(pr, E, pg, tm);
    const double result = *slk_mat_ptr(tm, pg.stream[4u * pg.op_start[hp.last_op] + 1u]);
    if(prob) {
        if(result <= 0.0) *prob = -DBL_MAX;
        else if(!dg) *prob = log(result);
        else {
            const int nmeioses = (sex_linked ? 1 : 2) * (N - F);
            const double recomb = (double) ncross * hp.log_theta[interval] + (double)(nmeioses - ncross) * hp.log_1mtheta[interval];
            *prob = log(result) - recomb - hp.marker_transmission;
        }
    }
    return result;
}

}  // extern "C"
