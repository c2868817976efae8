// slk_kernels.cuh -- sm_100a kernels for the peeling hot path: the block-Gibbs locus sampler
// (L-sampler) and LOD scoring, both driven by the flattened peel program of slk_types.h.  The
// per-thread work (forward tile, backward term, staging) lives in slk_peel.h.
//
// Execution model (B200: 148 SMs, 227 KB shared memory per CTA, FP64 on the CUDA cores):
//   * a TEAM of T threads (T = 32..512, chosen per plan) owns one unit of work at a time --
//     one marker locus for the sampler, one (interval, position) for LOD scoring -- and keeps
//     the peel matrices of that unit in its arena (shared memory; the few largest ones in an
//     L2-resident slab); persistent CTAs stride over the units, grid = SMs x resident CTAs;
//   * the program (pre-decoded op records, schedules) is one blob copied into shared memory once
//     per CTA;
//   * inside a unit the peel ops are walked level by level (dependency levels of the peel
//     forest); the rows of all ops of a level form one index space that the whole team strides
//     over, one 4 x 4 register tile (sixteen independent product chains) per thread, so a level
//     of many tiny ops costs one pass; imap names the op of an item with one load;
//   * only VALID rows are enumerated: the valid set of an op is a Cartesian product of
//     per-person legal genotype sets (4-bit masks), so the j-th valid row is a mixed-radix
//     decode of j -- no per-locus index lists (the reference's matrix_indices[locus] tables,
//     peel_sequence_generator.cc:141-159, are what stops it scaling to 10k loci);
//   * matrices are indexed by the cutset sorted by peel position, so a consumed function is read
//     as 16, 4 or 1 consecutive doubles per tile with vector loads (slk_types.h);
//   * the presum matrix is never stored: the backward pass recomputes the four candidates of
//     each op at the already sampled cutset assignment, four lanes per op, ops grouped into
//     backward levels; draws come from Philox keyed by (chain, iteration, locus, slot).
//
// Arithmetic follows the reference's operation order (compiled with -fmad=false) so peel
// matrices are bit-identical to the CPU R-functions: rfunction.cc:116-142,
// sampler_rfunction.cc:102-157,190-287,350-419, trait_rfunction.cc:9-137.
#ifndef SLK_KERNELS_CUH
#define SLK_KERNELS_CUH

// The kernels are compiled for a fixed list of (team threads, CTA threads) geometries (slk_geometry.h); the CTA
// size fixes the register budget of a thread (65536 / CTA threads, one CTA per SM).
#include <stdint.h>
#include <float.h>

#include "swiftlink_b200.h"
#include "slk_types.h"
#include "slk_philox.cuh"
#include "slk_peel.h"

struct SlkLaunch {
    SlkDevPlan plan;
    uint8_t* dgp;                // [M][N] descent graph, bit0 maternal, bit1 paternal
    double* lod;                 // [(M-1)*nlod] log-sum accumulators
    double* gscratch;            // global arena slab: teams x (arena_doubles - smem_doubles)
    int* err;                    // [0] first error code, [1] unit
    int* ticket;                 // sampler: next unit of the launch (zeroed by the host before the launch), or NULL
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


// ---- small device helpers ------------------------------------------------------------

template<int T>
__device__ __forceinline__ void team_sync(int team) {
    if(T == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" :: "r"(team + 1), "r"(T) : "memory");
}

// Brings the program blob into shared memory with ONE 1-D bulk copy (TMA: cp.async.bulk global -> shared, completion
// counted in bytes on an mbarrier) issued by one thread; every thread then waits on the barrier's phase.  Returns
// the shared views.  (blob and shared base are 16-byte aligned, blob_bytes is a multiple of 16.)
__device__ __forceinline__ SlkProgView stage_program(const SlkProgram& p, const double* dprob_global, unsigned char* smem) {
    // the barrier lives in the dynamic allocation, right behind the blob (static shared memory would count
    // against the 227 KB the dynamic part is allowed to ask for)
    const uint32_t bar = (uint32_t) __cvta_generic_to_shared(smem + p.blob_bytes);
    const uint32_t dst = (uint32_t) __cvta_generic_to_shared(smem);
    if(threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((uint32_t) p.blob_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst), "l"(p.blob), "r"((uint32_t) p.blob_bytes), "r"(bar) : "memory");
    }
    __syncthreads();                                       // the barrier is initialised before anybody polls it
    {
        uint32_t done = 0;
        while(!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(0u) : "memory");
        }
    }
    SlkProgView g;
    g.stream = (const uint32_t*)(smem + p.off_stream);
    g.op_start = (const uint16_t*)(smem + p.off_op_start);
    g.imap = (const uint16_t*)(smem + p.off_imap);
    g.flevel_items = (const uint32_t*)(smem + p.off_flevel_items);
    g.flevel_map = (const uint32_t*)(smem + p.off_flevel_map);
    g.bops = (const uint16_t*)(smem + p.off_bops);
    g.blevel_start = (const uint16_t*)(smem + p.off_blevel_start);
    g.glist = smem + p.off_glist;
    g.dprob = p.off_dprob >= 0 ? (const double*)(smem + p.off_dprob) : dprob_global;
    return g;
}

// Asynchronous 4-byte copies global -> shared (LDGSTS): the per-unit rows (genotype codes, neighbouring descent-graph
// rows; N bytes each, 4-byte aligned when N is a multiple of 4) are requested first and waited for after the unit's
// other staging work (Philox draws, priors), so their latency is behind that work.
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t) __cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ SlkTeam carve_team(const SlkLaunch& L, const SlkProgram& p, unsigned char* base, int gteam) {
    const SlkDevPlan& P = L.plan;
    SlkTeamLayout lay = slk_team_layout(P.N, P.F, P.nops, p.smem_doubles, p.table_doubles_per_child);
    SlkTeam tm;
    tm.arena = (double*)(base + lay.arena);
    tm.smem_doubles = p.smem_doubles;
    tm.garena = L.gscratch + (size_t) gteam * (size_t)(p.arena_doubles - p.smem_doubles);
    tm.tables = (double*)(base + lay.tables);
    tm.scal = (double*)(base + lay.scal);
    tm.oploc = (SlkOpLoc*)(base + lay.oploc);
    tm.ru = (double*)(base + lay.ru);
    tm.gc = base + lay.bytes;
    tm.dgl = tm.gc + lay.nbytes;
    tm.dgr = tm.dgl + lay.nbytes;
    tm.pmk = tm.dgr + lay.nbytes;
    tm.red = (int*)(base + lay.red);
    return tm;
}

__device__ __forceinline__ SlkEnv make_env(const SlkLaunch& L) {
    SlkEnv E;
    E.N = L.plan.N; E.F = L.plan.F; E.nops = L.plan.nops;
    E.sex_linked = L.plan.sex_linked;
    E.male = L.plan.male;
    E.dump_mat = L.dump_mat;
    E.dump_pre = L.dump_pre;
    E.dbg = L.plan.dbg;
    return E;
}

__device__ __forceinline__ void raise_error(int* err, int code, int unit) {
    if(atomicCAS(&err[0], 0, code) == 0) err[1] = unit;
}

// the forward pass of one unit: every level's items strided over the team
template<bool TRAIT, bool DBG, int T>
__device__ __forceinline__ void forward_levels(const SlkLaunch& L, const SlkEnv& E, const SlkProgram& pr, const SlkProgView& pg,
                                               const SlkTeam& tm, int team, int tid, bool tracing, int& tstamp) {
    for(int lv = 0; lv < pr.n_flevels; ++lv) {
        const uint32_t items = pg.flevel_items[lv];
        const uint32_t mw = pg.flevel_map[lv];
        const uint16_t* map = pg.imap + (mw & ~SLK_LEVEL_FINE);
        if(mw & SLK_LEVEL_FINE) {
            for(uint32_t q = tid; q < items; q += T) slk_forward_item<TRAIT, DBG>(E, pg, tm, map, true, q);
        }
        else {
            for(uint32_t q = tid; q < items; q += T) slk_forward_item<TRAIT, DBG>(E, pg, tm, map, false, q);
        }
        team_sync<T>(team);
        if(tracing) L.trace[tstamp++] = clock64();
    }
}

// ============================================================================================
// L-sampler: LocusSampler::set_locus_minimal + step (locus_sampler2.cc:128-181) for every locus
// of one window class.  Replaces lsampler_kernel (cuda_lsampler.cu:403-449).
// ============================================================================================
template<int T, int MAXCTA, bool DBG>
__global__ void __launch_bounds__(MAXCTA, 1) slk_lsampler_kernel(const SlkLaunch L) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SlkDevPlan& P = L.plan;
    const SlkProgram& pr = P.ls;
    const SlkProgView pg = stage_program(pr, P.person_prior, smem);      // dprob: the SLK_PRIOR_PERSON class's prior
    const SlkEnv E = make_env(L);

    const int teams_per_cta = blockDim.x / T;
    const int team = threadIdx.x / T;
    const int tid = threadIdx.x % T;
    const int lane = tid & 31;
    const int gteam = blockIdx.x * teams_per_cta + team;
    const int total_teams = gridDim.x * teams_per_cta;
    const SlkTeam tm = carve_team(L, pr, smem + pr.prog_smem_bytes + (size_t) team * pr.team_smem_bytes, gteam);
    const int N = P.N, F = P.F, M = P.M;

    // illegal peel genotypes and slots multiply stale cells by 0: the arena must never hold a NaN pattern
    for(int i = tid; i < pr.smem_doubles; i += T) tm.arena[i] = 0.0;

    // Batched sequential imputation (L.si_batch > 0): team g walks graph g -- M loci in sequence, rows [g M, (g + 1) M)
    // of the graph buffer, run number L.iteration + g, start locus L.si_starts[g] -- so that hundreds of the
    // reference's independent start_from runs (sequential_imputation.cc:47-115) share one launch.
    const int si_batch = L.si_batch;
    for(int graph = si_batch ? gteam : 0; graph < (si_batch ? si_batch : 1); graph += (si_batch ? total_teams : 1)) {
    uint8_t* const dgp = L.dgp + (si_batch ? (size_t) graph * (size_t) M * (size_t) N : (size_t) 0);
    const uint64_t iteration = L.iteration + (uint64_t)(si_batch ? graph : 0);
    const int si_start = si_batch ? L.si_starts[graph] : L.si_start;
    double* const si_weight = L.dump_result ? L.dump_result + (si_batch ? graph : 0) : (double*) 0;
    // One-warp teams take the next locus of the launch from a ticket counter (a team that drew cheap loci -- few legal
    // genotypes -- comes back sooner); larger teams and the sequential walks stride statically.
    const bool ticketed = T == 32 && L.ticket != 0 && !si_batch && si_start < 0;
    auto next_unit = [&](int prev) -> int {
        if(!ticketed) return prev + (si_batch ? 1 : total_teams);
        int u = 0;
        if(lane == 0) u = atomicAdd(L.ticket, 1);
        return __shfl_sync(0xffffffffu, u, 0);
    };
    for(int unit = ticketed ? next_unit(0) : (si_batch ? 0 : gteam); unit < L.nunits; unit = next_unit(unit)) {
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

        const bool tracing = L.trace != 0 && gteam == 0 && unit == gteam && tid == 0 && !ticketed;
        int tstamp = 0;
        if(tracing) L.trace[tstamp++] = clock64();
        // ---- stage: genotype codes, neighbouring descent-graph rows, thetas, founder priors
        if((N & 3) == 0) {
            // rows are 4-byte aligned: word loads (one 200-byte row is two or three sectors)
            const uint32_t* g4 = (const uint32_t*)(P.gcode + (size_t) plocus * N);
            const uint32_t* l4 = (const uint32_t*)(dgp + (size_t)(has_left ? locus - 1 : locus) * N);
            const uint32_t* r4 = (const uint32_t*)(dgp + (size_t)(has_right ? locus + 1 : locus) * N);
            for(int i = tid; i < (N >> 2); i += T) {
                cp_async4((uint32_t*) tm.gc + i, g4 + i);
                if(has_left) cp_async4((uint32_t*) tm.dgl + i, l4 + i); else ((uint32_t*) tm.dgl)[i] = 0u;
                if(has_right) cp_async4((uint32_t*) tm.dgr + i, r4 + i); else ((uint32_t*) tm.dgr)[i] = 0u;
            }
        }
        else {
            for(int i = tid; i < N; i += T) {
                tm.gc[i] = P.gcode[(size_t) plocus * N + i];
                tm.dgl[i] = has_left ? dgp[(size_t)(locus - 1) * N + i] : 0;
                tm.dgr[i] = has_right ? dgp[(size_t)(locus + 1) * N + i] : 0;
            }
        }
        for(int i = tid; i < 28; i += T)
            tm.scal[16 + i] = (i < 20) ? kSlkClassPrior[i >> 2][i & 3] : P.fprior[(size_t) plocus * 8 + (i - 20)];
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
        cp_async_wait_all();                               // the rows requested above (a no-op when none were)
        team_sync<T>(team);

        // per-child transmission tables and per-op legal masks of this locus
        for(int q = tid; q < 2 * (N - F); q += T) slk_stage_transmission(E, tm, q, has_left, has_right);
        for(int op = tid; op < P.nops; op += T) slk_stage_oploc<false>(pg, tm, op);
        team_sync<T>(team);

        if(tracing) L.trace[tstamp++] = clock64();
        // ---- forward peel, level by level
        forward_levels<false, DBG, T>(L, E, pr, pg, tm, team, tid, tracing, tstamp);

        const double result = *slk_mat_ptr(tm, pg.stream[4u * pg.op_start[P.last_op] + 1u]);
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
            // When the level has no more ops than the team has warps every op gets its own warp (lanes
            // 0-3 = the four candidate genotypes): ops of different shapes packed in one warp would
            // serialise each other's control flow on this latency-critical chain.
            const bool spread = (e - b) <= T / 32;
            for(int base = b; base < e; base += (spread ? T / 32 : T / 4)) {
                const int q = base + (spread ? (tid >> 5) : (tid >> 2));
                const uint32_t g = tid & 3u;
                double d = 0.0;
                int op = 0;
                const bool live = q < e && (!spread || lane < 4);
                if(live) {
                    op = pg.bops[q];
                    d = slk_backward_term(E, pg, tm, (uint32_t) op, g);
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
                    const int peel = pg.stream[4u * pg.op_start[op]] >> 16;
                    tm.pmk[peel] = (uint8_t) slk_sample4(dd, tm.ru[op]);
                }
            }
            team_sync<T>(team);
            if(tracing) L.trace[tstamp++] = clock64();
        }

        // ---- meiosis indicators (locus_sampler2.cc:32-125)
        {
            const bool use_left = has_left && !ign_left, use_right = has_right && !ign_right;
            const double th_left = use_left ? P.theta[plocus - 1] : 0.0, th_right = use_right ? P.theta[plocus] : 0.0;
            for(int i = F + tid; i < N; i += T) {
                const uint32_t out = slk_sample_indicators(E, tm, i, P.mother[i], P.father[i], use_left, use_right, th_left, th_right,
                    [&](int parent) { return slk_uniform(L.seed, L.chain, iteration, (uint32_t)(locus + L.row_base),
                                                         (uint32_t)(P.nops + 2 * i + parent)); });
                dgp[(size_t) locus * N + i] = (uint8_t) out;
            }
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
template<int T, int MAXCTA, bool DBG>
__global__ void __launch_bounds__(MAXCTA, 1) slk_lodscore_kernel(const SlkLaunch L) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SlkDevPlan& P = L.plan;
    const SlkProgram& pr = P.lod;
    const SlkProgView pg = stage_program(pr, P.disease_prob, smem);
    const SlkEnv E = make_env(L);

    const int teams_per_cta = blockDim.x / T;
    const int team = threadIdx.x / T;
    const int tid = threadIdx.x % T;
    const int lane = tid & 31;
    const int gteam = blockIdx.x * teams_per_cta + team;
    const int total_teams = gridDim.x * teams_per_cta;
    const SlkTeam tm = carve_team(L, pr, smem + pr.prog_smem_bytes + (size_t) team * pr.team_smem_bytes, gteam);
    const int N = P.N, F = P.F;
    const bool sex_linked = P.sex_linked != 0;
    const double trait_prob = sex_linked ? 0.5 : 0.25;                      // trait_rfunction.cc:37,78

    for(int i = tid; i < pr.smem_doubles; i += T) tm.arena[i] = 0.0;
    // the legal sets of the trait program do not depend on the position: staged once
    for(int op = tid; op < P.nops; op += T) slk_stage_oploc<true>(pg, tm, op);
    team_sync<T>(team);

    for(int u = gteam; u < L.nunits; u += total_teams) {
        const int unit = L.unit_base + u;
        const int interval = L.period ? 0 : unit / P.nlod;
        const int k = L.period ? 0 : unit - interval * P.nlod;              // position k+1 of nlod
        const size_t row_l = L.period ? (size_t) unit * L.period + L.lod_row0 : (size_t) interval;
        const size_t row_r = L.period ? (size_t) unit * L.period + L.lod_row1 : (size_t) interval + 1;

        int ncross = 0;
        if(!L.no_dg) {
            if((N & 3) == 0) {
                const uint32_t* l4 = (const uint32_t*)(L.dgp + row_l * N);
                const uint32_t* r4 = (const uint32_t*)(L.dgp + row_r * N);
                // (plain loads here: nothing follows that an asynchronous copy could hide behind; measured 12.25 against
                // 12.76 ms per pass with cp.async)
                for(int i = tid; i < (N >> 2); i += T) { ((uint32_t*) tm.dgl)[i] = l4[i]; ((uint32_t*) tm.dgr)[i] = r4[i]; }
            }
            else {
                for(int i = tid; i < N; i += T) {
                    tm.dgl[i] = L.dgp[row_l * N + i];
                    tm.dgr[i] = L.dgp[row_r * N + i];
                }
            }
            if(tid == 0) {
                // trait_rfunction.h:50-56
                const double th = P.partial[interval] * (double)(k + 1);
                const double th2 = P.partial[interval] * (double)(P.nlod + 1 - (k + 1));
                tm.scal[0] = th; tm.scal[1] = 1.0 - th; tm.scal[2] = th2; tm.scal[3] = 1.0 - th2;
            }
            team_sync<T>(team);
            // trait_prob x recombination probability per child and (i, j) (trait_rfunction.cc:9-22)
            for(int q = tid; q < 4 * (N - F); q += T) slk_stage_trait_weight(E, tm, q, trait_prob);
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

        int tstamp = 0;
        forward_levels<true, DBG, T>(L, E, pr, pg, tm, team, tid, false, tstamp);

        if(tid == 0) {
            const double result = *slk_mat_ptr(tm, pg.stream[4u * pg.op_start[P.last_op] + 1u]);
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
