// slk_capi.cu -- implementation of the C ABI in include/swiftlink_b200.h: plan upload, chain
// state, kernel launches.  There is deliberately no CPU fallback anywhere in this file: if no
// sm_100 device is usable every compute entry point fails with SLK_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <algorithm>
#include <vector>

#include "swiftlink_b200.h"
#include "slk_plan.h"
#include "slk_kernels.cuh"
#include "slk_geometry.h"
#include "slk_msampler.cuh"

namespace {

thread_local std::string g_error;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define CU(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) \
    return fail(SLK_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while(0)

template<typename T>
cudaError_t upload(const std::vector<T>& v, const T** out, std::vector<void*>& owned) {
    void* p = 0;
    size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    cudaError_t e = cudaMalloc(&p, bytes);
    if(e != cudaSuccess) return e;
    owned.push_back(p);
    if(!v.empty()) {
        e = cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
        if(e != cudaSuccess) return e;
    }
    *out = (const T*) p;
    return cudaSuccess;
}

}  // namespace

struct slk_plan {
    int device;
    int sm_count;
    slk::HostPlan host;
    SlkDevPlan dev;
    std::vector<void*> owned;
    int ls_blocks_per_sm, lod_blocks_per_sm;
    int ls_cta_smem, lod_cta_smem;
    // M-sampler likelihood kernel geometry
    int ms_grid, ms_threads, ms_smem;
    int ms_chain_loci, ms_chain_smem;
    int ms_step_smem;
};

struct slk_chain {
    slk_plan* plan;
    uint64_t seed;
    uint32_t chain_id;
    cudaStream_t own_stream, stream;
    uint8_t* dgp;
    int32_t* dg_staging;          // device int32[M][N][2] for upload/download conversion
    double* lod;
    double* lod_out;
    double* gscratch;
    size_t gscratch_doubles;
    int* err;                     // device [4]
    int32_t count;                // scoring passes (LODscores::count)
    // M-sampler state (allocated on first use)
    double* ms_lncur; double* ms_lnl; uint8_t* ms_bits; double* ms_fb; double* ms_out; void* ms_te; uint32_t* ms_stale;
    uint32_t* ms_snap;            // two forest snapshots of slk_ms_snap_words(F) x M words (slk_ms_step_kernel)
    bool ms_cur_valid;            // ms_cur describes the graph now in dgp
    unsigned long long* ms_timeline;   // slk_debug_msampler_timeline only: %globaltimer stamps of the sweep being traced
    int ms_tl_slot, ms_tl_cta_slot;
    // debug buffers (allocated on first use)
    double* dbg_mat; double* dbg_pre; double* dbg_dist4; int* dbg_pmk; double* dbg_res; double* dbg_prob;
};

namespace {

typedef void (*PeelKernel)(const SlkLaunch);
struct PeelGeom { int team, max_cta; PeelKernel ls, lod; };
#define SLK_X(T, C) { T, C, slk_lsampler_kernel<T, C, false>, slk_lodscore_kernel<T, C, false> },
const PeelGeom kGeoms[] = { SLK_GEOMETRIES(SLK_X) };
#undef SLK_X

// the instantiation for a team size with the smallest CTA bound that holds `cta` threads (most registers per thread)
const PeelGeom* find_geom(int team, int cta) {
    const PeelGeom* best = 0;
    for(size_t i = 0; i < sizeof(kGeoms) / sizeof(kGeoms[0]); ++i) {
        const PeelGeom& g = kGeoms[i];
        if(g.team == team && g.max_cta >= cta && (!best || g.max_cta < best->max_cta)) best = &g;
    }
    return best;
}

// parity hooks: one team of 128 threads whatever the plan's production geometry (the team size
// is a launch parameter only, the shared-memory layout does not depend on it)
const int kDebugTeam = 128;
cudaError_t prep_debug(int, int) {
    // the attribute belongs to the function, not to a plan: always allow the maximum so that
    // plans of different sizes can coexist in one process
    cudaError_t e = cudaFuncSetAttribute(slk_lsampler_kernel<kDebugTeam, 384, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, slk::kSmemPerBlockMax);
    if(e != cudaSuccess) return e;
    return cudaFuncSetAttribute(slk_lodscore_kernel<kDebugTeam, 384, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, slk::kSmemPerBlockMax);
}
void launch_ls_debug(const SlkLaunch& L, int smem, cudaStream_t s) {
    slk_lsampler_kernel<kDebugTeam, 384, true><<<1, kDebugTeam, smem, s>>>(L);
}
void launch_lod_debug(const SlkLaunch& L, int grid, int smem, cudaStream_t s) {
    slk_lodscore_kernel<kDebugTeam, 384, true><<<grid, kDebugTeam, smem, s>>>(L);
}

cudaError_t prep_by_team(int team, int ls_smem, int lod_smem, int ls_cta, int lod_cta, int* ls_occ, int* lod_occ, bool ls, bool lod) {
    cudaError_t e;
    if(ls) {
        const PeelGeom* g = find_geom(team, ls_cta);
        if(!g) return cudaErrorInvalidConfiguration;
        e = cudaFuncSetAttribute(g->ls, cudaFuncAttributeMaxDynamicSharedMemorySize, slk::kSmemPerBlockMax);
        if(e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(ls_occ, g->ls, ls_cta, ls_smem);
        if(e != cudaSuccess) return e;
    }
    if(lod) {
        const PeelGeom* g = find_geom(team, lod_cta);
        if(!g) return cudaErrorInvalidConfiguration;
        e = cudaFuncSetAttribute(g->lod, cudaFuncAttributeMaxDynamicSharedMemorySize, slk::kSmemPerBlockMax);
        if(e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(lod_occ, g->lod, lod_cta, lod_smem);
        if(e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// a single-team launch (sequential imputation, P(T)) uses the same instantiation as the full CTA
void launch_ls_by_team(int team, const SlkLaunch& L, int grid, int cta, int smem, cudaStream_t s) {
    const PeelGeom* g = find_geom(team, std::max(cta, L.plan.ls.cta_threads));
    g->ls<<<grid, cta, smem, s>>>(L);
}

void launch_lod_by_team(int team, const SlkLaunch& L, int grid, int cta, int smem, cudaStream_t s) {
    const PeelGeom* g = find_geom(team, std::max(cta, L.plan.lod.cta_threads));
    g->lod<<<grid, cta, smem, s>>>(L);
}

int grid_for(int nunits, int teams_per_cta, int blocks_per_sm, int sm_count) {
    int need = (nunits + teams_per_cta - 1) / teams_per_cta;
    int cap = blocks_per_sm * sm_count;
    int g = need < cap ? need : cap;
    return g < 1 ? 1 : g;
}

SlkLaunch base_launch(slk_chain* c) {
    SlkLaunch L;
    memset(&L, 0, sizeof(L));
    L.plan = c->plan->dev;
    L.dgp = c->dgp;
    L.lod = c->lod;
    L.gscratch = c->gscratch;
    L.err = c->err;
    L.seed = c->seed;
    L.chain = c->chain_id;
    L.window = 2;
    L.dump_k = -1;
    L.si_start = -1;
    return L;
}

int check_device_error(slk_chain* c) {
    int h[2] = {0, 0};
    cudaError_t e = cudaMemcpyAsync(h, c->err, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
    if(e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if(e != cudaSuccess) return fail(SLK_ERR_CUDA, "device error: %s", cudaGetErrorString(e));
    if(h[0] != 0) {
        cudaMemsetAsync(c->err, 0, 4 * sizeof(int), c->stream);
        if(h[0] == SLK_ERR_ZERO_LIKELIHOOD)
            return fail(h[0], "likelihood is zero at locus %d (check penetrance function?)", h[1]);
        if(h[0] == SLK_ERR_ILLEGAL_GRAPH)
            return fail(h[0], "illegal descent graph given to m-sampler (locus %d)", h[1]);
        return fail(h[0], "intermediate state had a likelihood of 0.0 or less (lod score unit %d)", h[1]);
    }
    return SLK_OK;
}

int ensure_debug(slk_chain* c) {
    if(c->dbg_mat) return SLK_OK;
    const slk::HostPlan& hp = c->plan->host;
    CU(cudaMalloc((void**) &c->dbg_mat, sizeof(double) * hp.sum_cells));
    CU(cudaMalloc((void**) &c->dbg_pre, sizeof(double) * hp.sum_presum));
    CU(cudaMalloc((void**) &c->dbg_dist4, sizeof(double) * 4 * hp.nops));
    CU(cudaMalloc((void**) &c->dbg_pmk, sizeof(int) * hp.N));
    CU(cudaMalloc((void**) &c->dbg_res, sizeof(double) * (hp.nlod + 1)));
    CU(cudaMalloc((void**) &c->dbg_prob, sizeof(double) * (hp.nlod + 1)));
    return SLK_OK;
}

}  // namespace

extern "C" {

int slk_abi_version(void) { return SLK_ABI_VERSION; }

const char* slk_last_error(void) { return g_error.c_str(); }

int slk_device_count(void) {
    int n = 0;
    if(cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int slk_plan_create(const slk_problem* problem, int device, slk_plan** out) {
    if(!problem || !out) return fail(SLK_ERR_INVALID, "null argument");
    *out = 0;
    slk_plan* p = new slk_plan();
    std::string err;
    if(!slk::build_plan(*problem, p->host, err)) {
        delete p;
        return fail(SLK_ERR_INVALID, "%s", err.c_str());
    }
    int ndev = slk_device_count();
    if(ndev == 0 || device < 0 || device >= ndev) {
        delete p;
        return fail(SLK_ERR_NO_DEVICE, "no usable CUDA device %d (found %d): this library has no CPU path", device, ndev);
    }
    p->device = device;
    cudaDeviceProp prop;
    if(cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete p;
        return fail(SLK_ERR_CUDA, "cannot select device %d", device);
    }
    if(prop.major < 10) {
        delete p;
        return fail(SLK_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    }
    p->sm_count = prop.multiProcessorCount;

    const slk::HostPlan& h = p->host;
    SlkDevPlan& d = p->dev;
    memset(&d, 0, sizeof(d));
    d.N = h.N; d.F = h.F; d.M = h.M; d.nlod = h.nlod; d.sex_linked = h.sex_linked; d.nops = h.nops; d.last_op = h.last_op;
    d.marker_transmission = h.marker_transmission;
    cudaError_t e = cudaSuccess;
#define UP(field, vec) if(e == cudaSuccess) e = upload(vec, &d.field, p->owned)
    UP(mother, h.mother); UP(father, h.father); UP(male, h.male); UP(disease_prob, h.disease_prob); UP(person_prior, h.person_prior);
    UP(gcode, h.gcode); UP(fprior, h.fprior); UP(theta, h.theta); UP(partial, h.partial);
    UP(log_theta, h.log_theta); UP(log_1mtheta, h.log_1mtheta);
    for(int k = 0; k < 2; ++k) {
        const slk::HostProgram& hp = k ? h.lod : h.ls;
        SlkProgram& dp = k ? d.lod : d.ls;
        dp = hp.layout;
        if(e == cudaSuccess) e = upload(hp.blob, &dp.blob, p->owned);
    }
    UP(dbg, h.ls.dbg);
    UP(ms.seq, h.ms_seq); UP(ms.typed, h.ms_typed); UP(ms.obsT, h.ms_obsT); UP(ms.minor, h.ms_minor);
    UP(ms.lnmajor, h.ms_lnmajor); UP(ms.lnminor, h.ms_lnminor);
    UP(ms.desc_mask, h.ms_desc_mask); UP(ms.typed_index, h.ms_typed_index);
    d.ms.W = h.ms_W;
    d.ms.n_typed = (int) h.ms_typed.size();
    d.ms.available = h.ms_available ? 1 : 0;
#undef UP
    if(e != cudaSuccess) {
        std::string msg = cudaGetErrorString(e);
        slk_plan_destroy(p);
        return fail(SLK_ERR_CUDA, "plan upload failed: %s", msg.c_str());
    }
    {
        // M-sampler likelihood kernel: one thread per locus, loci dealt evenly over at most one CTA per
        // SM; as many threads as the per-thread tables leave room for in shared memory
        const MsLayout lay = slk_ms_layout(h.N, h.F, d.ms.n_typed);
        // one warp per CTA (the warps are independent): small CTAs let several chains' launches share
        // an SM, and every locus of a 10k-marker map is resident at once (3 CTAs per SM)
        p->ms_threads = 0;
        if((int) lay.cta_tables + 32 * (int) lay.per_thread <= slk::kSmemPerBlockMax) {
            p->ms_grid = (h.M + 31) / 32; p->ms_threads = 32;
            p->ms_smem = (int) lay.cta_tables + 32 * (int) lay.per_thread;
        }
        if(lay.G > SLK_MS_MAXGROUPS || h.N >= 2048) p->ms_threads = 0;
        {
            const MsStepLayout sl = slk_ms_step_layout(h.N, h.F, d.ms.n_typed, h.ms_W);
            p->ms_step_smem = (int) sl.cta_tables + 32 * (int) sl.per_thread;
        }      // 12-bit component ids, 16-bit row offsets
        // chain kernel: raw pair (16 B) + one map byte per locus of the CTA's share of the chromosome
        {
            const int per_thread = (h.M + SLK_MS_CLUSTER * SLK_MS_CHAIN_THREADS - 1) / (SLK_MS_CLUSTER * SLK_MS_CHAIN_THREADS);
            p->ms_chain_loci = per_thread * SLK_MS_CHAIN_THREADS;
            p->ms_chain_smem = (p->ms_chain_loci * SLK_MS_CHAIN_BYTES_PER_LOCUS + 15) & ~15;
            if(p->ms_chain_smem > slk::kSmemPerBlockMax - 4096) p->ms_threads = 0;     // > 100k markers: unsupported for now
        }
    }

    p->ls_cta_smem = d.ls.prog_smem_bytes + (d.ls.cta_threads / d.ls.team_threads) * d.ls.team_smem_bytes;
    p->lod_cta_smem = d.lod.prog_smem_bytes + (d.lod.cta_threads / d.lod.team_threads) * d.lod.team_smem_bytes;
    if(p->ls_cta_smem > slk::kSmemPerBlockMax || p->lod_cta_smem > slk::kSmemPerBlockMax) {
        slk_plan_destroy(p);
        return fail(SLK_ERR_UNSUPPORTED, "peel program needs %d / %d bytes of shared memory per CTA", p->ls_cta_smem, p->lod_cta_smem);
    }
    p->ls_blocks_per_sm = p->lod_blocks_per_sm = 1;
    e = prep_by_team(d.ls.team_threads, p->ls_cta_smem, 0, d.ls.cta_threads, 0, &p->ls_blocks_per_sm, 0, true, false);
    if(e == cudaSuccess)
        e = prep_by_team(d.lod.team_threads, 0, p->lod_cta_smem, 0, d.lod.cta_threads, 0, &p->lod_blocks_per_sm, false, true);
    if(e == cudaSuccess)
        e = prep_debug(d.ls.prog_smem_bytes + d.ls.team_smem_bytes, d.lod.prog_smem_bytes + d.lod.team_smem_bytes);
    if(e == cudaSuccess) e = cudaFuncSetAttribute(slk_ms_likelihood_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, slk::kSmemPerBlockMax);
    if(e == cudaSuccess) e = cudaFuncSetAttribute(slk_ms_likelihood_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, slk::kSmemPerBlockMax);
    if(e == cudaSuccess) e = cudaFuncSetAttribute(slk_ms_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, slk::kSmemPerBlockMax);
    if(e == cudaSuccess) e = cudaFuncSetAttribute(slk_ms_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, slk::kSmemPerBlockMax);
    if(e == cudaSuccess) e = cudaFuncSetAttribute(slk_ms_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, slk::kSmemPerBlockMax - 4096);
    if(e != cudaSuccess || p->ls_blocks_per_sm < 1 || p->lod_blocks_per_sm < 1) {
        std::string msg = cudaGetErrorString(e);
        slk_plan_destroy(p);
        return fail(SLK_ERR_CUDA, "kernel configuration failed: %s", msg.c_str());
    }
    *out = p;
    return SLK_OK;
}

void slk_plan_destroy(slk_plan* p) {
    if(!p) return;
    for(size_t i = 0; i < p->owned.size(); ++i) cudaFree(p->owned[i]);
    delete p;
}

static int fill_stats(const slk::HostPlan& h, const slk_plan* p, double* out, int cap) {
    double v[20] = {
        (double) h.nops, (double) h.sum_cells, (double) h.sum_presum, h.flops_ls, h.flops_lod,
        (double) h.ls.flevel_items.size(), (double)(h.ls.blevel_start.size() - 1),
        (double) h.lod.flevel_items.size(), (double) h.ls.arena_doubles, (double) h.lod.arena_doubles,
        (double) h.lod.valid_cells, (double) h.max_cutset, (double) h.ls.team_threads, (double) h.lod.team_threads,
        (double) h.ls.smem_doubles, (double) h.lod.smem_doubles,
        p ? (double) p->ls_blocks_per_sm : 0.0, p ? (double) p->lod_blocks_per_sm : 0.0,
        p ? (double) p->ls_cta_smem : 0.0, p ? (double) p->lod_cta_smem : 0.0
    };
    int n = cap < 20 ? cap : 20;
    for(int i = 0; i < n; ++i) out[i] = v[i];
    return n;
}

int slk_plan_stats(const slk_plan* p, double* out, int cap) {
    if(!p || !out) return 0;
    return fill_stats(p->host, p, out, cap);
}

int slk_plan_validate(const slk_problem* problem, double* stats, int cap) {
    if(!problem) return fail(SLK_ERR_INVALID, "null argument");
    slk::HostPlan hp;
    std::string err;
    if(!slk::build_plan(*problem, hp, err)) return fail(SLK_ERR_INVALID, "%s", err.c_str());
    if(stats) fill_stats(hp, 0, stats, cap);
    return SLK_OK;
}

// The global part of the peel arenas is scratch: every unit overwrites its team's slice and nothing
// outlives a launch.  Marking the slab as persisting in L2 keeps the slices resident between units, so
// the dirty lines are overwritten in place instead of being written back to HBM (ncu on the 200 x 10k
// workload: 352 MB of DRAM writes per L-sampler launch without the window).
static void pin_scratch_in_l2(slk_chain* c, cudaStream_t s) {
    if(getenv("SLK_NO_L2_WINDOW") || c->gscratch_doubles < 1024) return;
    cudaDeviceProp prop;
    if(cudaGetDeviceProperties(&prop, c->plan->device) != cudaSuccess) { cudaGetLastError(); return; }
    size_t bytes = c->gscratch_doubles * sizeof(double);
    size_t carve = std::min<size_t>(bytes, (size_t) prop.persistingL2CacheMaxSize);
    if(carve == 0 || prop.accessPolicyMaxWindowSize <= 0) return;
    if(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) != cudaSuccess) { cudaGetLastError(); return; }
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.base_ptr = c->gscratch;
    attr.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t) prop.accessPolicyMaxWindowSize);
    attr.accessPolicyWindow.hitRatio = (float) std::min(1.0, (double) carve / (double) attr.accessPolicyWindow.num_bytes);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
}

int slk_chain_create(slk_plan* plan, uint64_t seed, uint32_t chain_id, slk_chain** out) {
    if(!plan || !out) return fail(SLK_ERR_INVALID, "null argument");
    *out = 0;
    CU(cudaSetDevice(plan->device));
    slk_chain* c = new slk_chain();
    memset(c, 0, sizeof(*c));
    c->plan = plan; c->seed = seed; c->chain_id = chain_id;
    const SlkDevPlan& d = plan->dev;
    const size_t cells = (size_t) d.M * d.N;
    const int nlod_total = (d.M - 1) * d.nlod;
    // global arena slab: one slice per resident team of either kernel
    size_t ls_teams = (size_t) plan->ls_blocks_per_sm * plan->sm_count * (d.ls.cta_threads / d.ls.team_threads);
    size_t lod_teams = (size_t) plan->lod_blocks_per_sm * plan->sm_count * (d.lod.cta_threads / d.lod.team_threads);
    size_t sc = std::max(ls_teams * (size_t)(d.ls.arena_doubles - d.ls.smem_doubles),
                         lod_teams * (size_t)(d.lod.arena_doubles - d.lod.smem_doubles));
    c->gscratch_doubles = sc;
    cudaError_t e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    c->stream = c->own_stream;
    if(e == cudaSuccess) e = cudaMalloc((void**) &c->dgp, cells);
    if(e == cudaSuccess) e = cudaMalloc((void**) &c->dg_staging, cells * 2 * sizeof(int32_t));
    if(e == cudaSuccess) e = cudaMalloc((void**) &c->lod, sizeof(double) * nlod_total);
    if(e == cudaSuccess) e = cudaMalloc((void**) &c->lod_out, sizeof(double) * nlod_total);
    if(e == cudaSuccess) e = cudaMalloc((void**) &c->gscratch, sizeof(double) * std::max<size_t>(sc, 2));
    if(e == cudaSuccess) e = cudaMalloc((void**) &c->err, 4 * sizeof(int));
    if(e == cudaSuccess) e = cudaMemset(c->err, 0, 4 * sizeof(int));
    if(e == cudaSuccess) e = cudaMemset(c->dgp, 0, cells);
    if(e != cudaSuccess) {
        std::string msg = cudaGetErrorString(e);
        slk_chain_destroy(c);
        return fail(SLK_ERR_CUDA, "chain allocation failed: %s", msg.c_str());
    }
    pin_scratch_in_l2(c, c->own_stream);
    *out = c;
    return slk_lodscore_init(c);
}

void slk_chain_destroy(slk_chain* c) {
    if(!c) return;
    cudaSetDevice(c->plan->device);
    if(c->own_stream) { cudaStreamSynchronize(c->own_stream); cudaStreamDestroy(c->own_stream); }
    cudaFree(c->dgp); cudaFree(c->dg_staging); cudaFree(c->lod); cudaFree(c->lod_out);
    cudaFree(c->gscratch); cudaFree(c->err);
    cudaFree(c->dbg_mat); cudaFree(c->dbg_pre); cudaFree(c->dbg_dist4); cudaFree(c->dbg_pmk);
    cudaFree(c->dbg_res); cudaFree(c->dbg_prob);
    cudaFree(c->ms_lncur); cudaFree(c->ms_lnl); cudaFree(c->ms_bits); cudaFree(c->ms_te); cudaFree(c->ms_stale); cudaFree(c->ms_snap); cudaFree(c->ms_fb); cudaFree(c->ms_out);
    delete c;
}

int slk_chain_set_stream(slk_chain* c, void* cuda_stream) {
    if(!c) return fail(SLK_ERR_INVALID, "null chain");
    CU(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t) cuda_stream : c->own_stream;
    if(cuda_stream) pin_scratch_in_l2(c, c->stream);
    return SLK_OK;
}

int slk_chain_sync(slk_chain* c) {
    if(!c) return fail(SLK_ERR_INVALID, "null chain");
    CU(cudaSetDevice(c->plan->device));
    return check_device_error(c);
}

int slk_dg_upload(slk_chain* c, const int32_t* dg) {
    if(!c || !dg) return fail(SLK_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->plan->device));
    const size_t cells = (size_t) c->plan->dev.M * c->plan->dev.N;
    CU(cudaMemcpyAsync(c->dg_staging, dg, cells * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    slk_dg_pack_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, c->stream>>>(c->dg_staging, c->dgp, cells);
    CU(cudaGetLastError());
    c->ms_cur_valid = false;
    // the caller may reuse its buffer as soon as we return (GPULodscores::calculate uses a
    // synchronous copy for the same reason, gpu_lodscores.cc:598-607)
    CU(cudaStreamSynchronize(c->stream));
    return SLK_OK;
}

int slk_dg_download(slk_chain* c, int32_t* dg) {
    if(!c || !dg) return fail(SLK_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->plan->device));
    const size_t cells = (size_t) c->plan->dev.M * c->plan->dev.N;
    slk_dg_unpack_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, c->stream>>>(c->dgp, c->dg_staging, cells);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(dg, c->dg_staging, cells * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    return check_device_error(c);
}

int slk_dg_swap(slk_chain* a, slk_chain* b) {
    if(!a || !b) return fail(SLK_ERR_INVALID, "null chain");
    const SlkDevPlan& da = a->plan->dev;
    const SlkDevPlan& db = b->plan->dev;
    if(a->plan->device != b->plan->device || da.M != db.M || da.N != db.N)
        return fail(SLK_ERR_INVALID, "slk_dg_swap: chains must live on one device and have the same dimensions");
    CU(cudaSetDevice(a->plan->device));
    CU(cudaStreamSynchronize(a->stream));
    CU(cudaStreamSynchronize(b->stream));
    std::swap(a->dgp, b->dgp);
    a->ms_cur_valid = b->ms_cur_valid = false;     // the carried likelihoods belong to the other plan's frequencies
    return SLK_OK;
}

int slk_lsampler_window(slk_chain* c, uint64_t iteration, int window_length, int offset) {
    if(!c) return fail(SLK_ERR_INVALID, "null chain");
    if(window_length < 2 || offset < 0 || offset >= window_length)
        return fail(SLK_ERR_INVALID, "window_length must be >= 2 and 0 <= offset < window_length");
    slk_plan* p = c->plan;
    const SlkDevPlan& d = p->dev;
    CU(cudaSetDevice(p->device));
    SlkLaunch L = base_launch(c);
    L.iteration = iteration;
    L.window = window_length;
    L.offset = offset;
    L.nunits = (d.M - offset + window_length - 1) / window_length;
    if(L.nunits <= 0) return SLK_OK;
    c->ms_cur_valid = false;
    const int tpc = d.ls.cta_threads / d.ls.team_threads;
    const int grid = grid_for(L.nunits, tpc, p->ls_blocks_per_sm, p->sm_count);
    static const bool no_ticket = getenv("SLK_LS_NO_TICKET") != 0;          // tuning aid
    if(d.ls.team_threads == 32 && !no_ticket) {
        L.ticket = c->err + 2;                                                // err[2]: the launch's unit counter
        CU(cudaMemsetAsync(L.ticket, 0, sizeof(int), c->stream));
    }
    launch_ls_by_team(d.ls.team_threads, L, grid, d.ls.cta_threads, p->ls_cta_smem, c->stream);
    CU(cudaGetLastError());
    return SLK_OK;
}

int slk_lsampler_locus_by_locus(slk_chain* c, uint64_t iteration) {
    if(!c) return fail(SLK_ERR_INVALID, "null chain");
    slk_plan* p = c->plan;
    const SlkDevPlan& d = p->dev;
    CU(cudaSetDevice(p->device));
    SlkLaunch L = base_launch(c);
    L.iteration = iteration;
    L.window = 1;
    L.offset = 0;
    L.nunits = d.M;
    L.ignore_left = L.ignore_right = 1;
    c->ms_cur_valid = false;
    const int tpc = d.ls.cta_threads / d.ls.team_threads;
    const int grid = grid_for(L.nunits, tpc, p->ls_blocks_per_sm, p->sm_count);
    launch_ls_by_team(d.ls.team_threads, L, grid, d.ls.cta_threads, p->ls_cta_smem, c->stream);
    CU(cudaGetLastError());
    return SLK_OK;
}

int slk_sequential_imputation(slk_chain* c, uint64_t run, int start_locus, double* log_weight) {
    if(!c) return fail(SLK_ERR_INVALID, "null chain");
    slk_plan* p = c->plan;
    const SlkDevPlan& d = p->dev;
    if(start_locus < 0 || start_locus >= d.M) return fail(SLK_ERR_INVALID, "start locus out of range");
    CU(cudaSetDevice(p->device));
    int rc = ensure_debug(c);
    if(rc != SLK_OK) return rc;
    SlkLaunch L = base_launch(c);
    L.iteration = run;
    L.window = 1;
    L.offset = 0;
    L.nunits = d.M;
    L.si_start = start_locus;
    L.dump_result = c->dbg_res;
    c->ms_cur_valid = false;
    // one team walks the loci in sequence: a single CTA whose team 0 owns every unit
    launch_ls_by_team(d.ls.team_threads, L, 1, d.ls.team_threads, d.ls.prog_smem_bytes + d.ls.team_smem_bytes, c->stream);
    CU(cudaGetLastError());
    if(log_weight) {
        CU(cudaMemcpyAsync(log_weight, c->dbg_res, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        return check_device_error(c);
    }
    return SLK_OK;
}

// SequentialImputation::parallel_run (sequential_imputation.cc:47-115): `n_runs` independent start_from walks
// (run numbers first_run .. first_run + n_runs - 1, start locus start_loci[i]), one team each, in waves of as many
// teams as are resident; the best (first maximal) weight's graph becomes the chain's graph.  Same draws, weights
// and graphs as n_runs calls of slk_sequential_imputation, which uses one team of one SM at a time.
int slk_sequential_imputation_batch(slk_chain* c, uint64_t first_run, int n_runs, const int32_t* start_loci,
                                    double* log_weights, int32_t* best_run) {
    if(!c || !start_loci || n_runs < 1) return fail(SLK_ERR_INVALID, "null argument");
    slk_plan* p = c->plan;
    const SlkDevPlan& d = p->dev;
    for(int i = 0; i < n_runs; ++i)
        if(start_loci[i] < 0 || start_loci[i] >= d.M) return fail(SLK_ERR_INVALID, "start locus out of range");
    CU(cudaSetDevice(p->device));
    const size_t cells = (size_t) d.M * d.N;
    const int tpc = d.ls.cta_threads / d.ls.team_threads;
    const int teams = p->ls_blocks_per_sm * p->sm_count * tpc;        // one wave: the scratch slab has a slice per resident team
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    int wave = std::min(n_runs, teams);
    wave = (int) std::max<size_t>(1, std::min<size_t>((size_t) wave, (free_b / 2) / std::max<size_t>(cells, 1)));
    uint8_t* graphs = 0; int* starts = 0; double* weights = 0;
    CU(cudaMalloc((void**) &graphs, cells * (size_t) wave));
    cudaError_t e = cudaMalloc((void**) &starts, sizeof(int) * (size_t) wave);
    if(e == cudaSuccess) e = cudaMalloc((void**) &weights, sizeof(double) * (size_t) wave);
    int rc = e == cudaSuccess ? SLK_OK : fail(SLK_ERR_CUDA, "sequential imputation buffers: %s", cudaGetErrorString(e));
    std::vector<double> w((size_t) wave);
    double best = 0.0; int best_i = -1;
    c->ms_cur_valid = false;
    for(int r0 = 0; r0 < n_runs && rc == SLK_OK; r0 += wave) {
        const int n = std::min(wave, n_runs - r0);
        e = cudaMemsetAsync(graphs, 0, cells * (size_t) n, c->stream);
        if(e == cudaSuccess) e = cudaMemcpyAsync(starts, start_loci + r0, sizeof(int) * (size_t) n, cudaMemcpyHostToDevice, c->stream);
        if(e != cudaSuccess) { rc = fail(SLK_ERR_CUDA, "sequential imputation: %s", cudaGetErrorString(e)); break; }
        SlkLaunch L = base_launch(c);
        L.dgp = graphs;
        L.iteration = first_run + (uint64_t) r0;
        L.window = 1; L.offset = 0; L.nunits = d.M;
        L.si_start = 0; L.si_batch = n; L.si_starts = starts;
        L.dump_result = weights;
        const int grid = std::min((n + tpc - 1) / tpc, p->ls_blocks_per_sm * p->sm_count);
        launch_ls_by_team(d.ls.team_threads, L, grid, d.ls.cta_threads, p->ls_cta_smem, c->stream);
        e = cudaGetLastError();
        if(e == cudaSuccess) e = cudaMemcpyAsync(w.data(), weights, sizeof(double) * (size_t) n, cudaMemcpyDeviceToHost, c->stream);
        if(e != cudaSuccess) { rc = fail(SLK_ERR_CUDA, "sequential imputation: %s", cudaGetErrorString(e)); break; }
        rc = check_device_error(c);
        if(rc != SLK_OK) break;
        int wave_best = -1;
        for(int i = 0; i < n; ++i) {
            if(log_weights) log_weights[r0 + i] = w[i];
            if(best_i < 0 || w[i] > best) { best = w[i]; best_i = r0 + i; wave_best = i; }
        }
        if(wave_best >= 0) {
            e = cudaMemcpyAsync(c->dgp, graphs + cells * (size_t) wave_best, cells, cudaMemcpyDeviceToDevice, c->stream);
            if(e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if(e != cudaSuccess) rc = fail(SLK_ERR_CUDA, "sequential imputation: %s", cudaGetErrorString(e));
        }
    }
    cudaFree(graphs); cudaFree(starts); cudaFree(weights);
    if(best_run) *best_run = best_i;
    return rc;
}

int slk_lsampler_sweep(slk_chain* c, uint64_t iteration) {
    if(!c) return fail(SLK_ERR_INVALID, "null chain");
    // same draw as the oracle: first parity class of the sweep
    const int first = slk_uniform(c->seed, c->chain_id, iteration, 0u, SLK_SLOT_PHASE) < 0.5 ? 0 : 1;
    int rc = slk_lsampler_window(c, iteration, 2, first);
    if(rc != SLK_OK) return rc;
    return slk_lsampler_window(c, iteration, 2, 1 - first);
}

int slk_lodscore_init(slk_chain* c) {
    if(!c) return fail(SLK_ERR_INVALID, "null chain");
    CU(cudaSetDevice(c->plan->device));
    const int n = (c->plan->dev.M - 1) * c->plan->dev.nlod;
    slk_lod_init_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->lod, n);
    CU(cudaGetLastError());
    c->count = 0;
    return SLK_OK;
}

int slk_lodscore_accumulate(slk_chain* c) {
    if(!c) return fail(SLK_ERR_INVALID, "null chain");
    slk_plan* p = c->plan;
    const SlkDevPlan& d = p->dev;
    CU(cudaSetDevice(p->device));
    SlkLaunch L = base_launch(c);
    L.accumulate = 1;
    L.nunits = (d.M - 1) * d.nlod;
    const int tpc = d.lod.cta_threads / d.lod.team_threads;
    const int grid = grid_for(L.nunits, tpc, p->lod_blocks_per_sm, p->sm_count);
    launch_lod_by_team(d.lod.team_threads, L, grid, d.lod.cta_threads, p->lod_cta_smem, c->stream);
    CU(cudaGetLastError());
    c->count += 1;
    return SLK_OK;
}

int slk_lodscore_read(slk_chain* c, double* raw, int32_t* count) {
    if(!c) return fail(SLK_ERR_INVALID, "null chain");
    CU(cudaSetDevice(c->plan->device));
    const int n = (c->plan->dev.M - 1) * c->plan->dev.nlod;
    if(raw) CU(cudaMemcpyAsync(raw, c->lod, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    if(count) *count = c->count;
    return check_device_error(c);
}

int slk_lodscore_normalise(slk_chain* c, double trait_prob, double* out) {
    if(!c || !out) return fail(SLK_ERR_INVALID, "null argument");
    if(c->count < 1) return fail(SLK_ERR_INVALID, "no scoring pass has been accumulated");
    CU(cudaSetDevice(c->plan->device));
    const int n = (c->plan->dev.M - 1) * c->plan->dev.nlod;
    slk_lod_normalise_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->lod, c->lod_out, n, log((double) c->count), trait_prob);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, c->lod_out, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    return check_device_error(c);
}

int slk_trait_likelihood(slk_plan* plan, double* log_prob) {
    if(!plan || !log_prob) return fail(SLK_ERR_INVALID, "null argument");
    slk_chain* c = 0;
    int rc = slk_chain_create(plan, 0, 0, &c);
    if(rc != SLK_OK) return rc;
    rc = ensure_debug(c);
    if(rc == SLK_OK) {
        const SlkDevPlan& d = plan->dev;
        SlkLaunch L = base_launch(c);
        L.no_dg = 1;
        L.accumulate = 0;
        L.nunits = 1;
        L.dump_result = c->dbg_res;
        L.dump_prob = c->dbg_prob;
        launch_lod_by_team(d.lod.team_threads, L, 1, d.lod.cta_threads, plan->lod_cta_smem, c->stream);
        cudaError_t e = cudaGetLastError();
        if(e == cudaSuccess) e = cudaMemcpyAsync(log_prob, c->dbg_prob, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if(e != cudaSuccess) rc = fail(SLK_ERR_CUDA, "trait likelihood launch failed: %s", cudaGetErrorString(e));
        else rc = check_device_error(c);
    }
    slk_chain_destroy(c);
    return rc;
}

// ---- M-sampler ---------------------------------------------------------------------------

static int ms_ready(slk_chain* c) {
    if(!c) return fail(SLK_ERR_INVALID, "null chain");
    slk_plan* p = c->plan;
    if(!p->dev.ms.available) return fail(SLK_ERR_UNSUPPORTED, "slk_problem.minor_freq was NULL: the M-sampler tables were not built");
    if(p->ms_threads < 32)
        return fail(SLK_ERR_UNSUPPORTED, "pedigree or map too large for the M-sampler's shared-memory tables (N=%d, F=%d, M=%d)", p->dev.N, p->dev.F, p->dev.M);
    CU(cudaSetDevice(p->device));
    if(!c->ms_lncur) {
        const size_t M = (size_t) p->dev.M;
        CU(cudaMalloc((void**) &c->ms_lncur, sizeof(double) * M));
        CU(cudaMalloc((void**) &c->ms_lnl, sizeof(double) * SLK_MS_MAXSETS * M));
        CU(cudaMalloc((void**) &c->ms_bits, 2 * M));
        CU(cudaMalloc((void**) &c->ms_fb, sizeof(double) * 2 * M));
        CU(cudaMalloc((void**) &c->ms_out, sizeof(double) * 4));
        CU(cudaMalloc((void**) &c->ms_te, (size_t) std::max(p->dev.ms.n_typed, 1) * M * 4));
        CU(cudaMalloc((void**) &c->ms_stale, sizeof(uint32_t) * M * p->dev.ms.W));
        CU(cudaMalloc((void**) &c->ms_snap, sizeof(uint32_t) * 2 * (size_t) slk_ms_snap_words(p->dev.F) * M));
        CU(cudaMemsetAsync(c->ms_stale, 0, sizeof(uint32_t) * M * p->dev.ms.W, c->stream));
        CU(cudaMemsetAsync(c->ms_fb, 0, sizeof(double) * 2 * M, c->stream));
        c->ms_cur_valid = false;
    }
    return SLK_OK;
}

static SlkMsLaunch ms_launch(slk_chain* c) {
    const SlkDevPlan& d = c->plan->dev;
    SlkMsLaunch L;
    memset(&L, 0, sizeof(L));
    L.ms = d.ms;
    L.N = d.N; L.F = d.F; L.M = d.M; L.sex_linked = d.sex_linked;
    L.mother = d.mother; L.father = d.father; L.male = d.male; L.theta = d.theta;
    L.log_theta = d.log_theta; L.log_1mtheta = d.log_1mtheta;
    L.dgp = c->dgp; L.lncur = c->ms_lncur; L.lnl = c->ms_lnl; L.bits = c->ms_bits; L.fb = c->ms_fb; L.err = c->err;
    L.te = c->ms_te; L.stale = c->ms_stale;
    L.snap = c->ms_snap; L.snap_use = -1; L.snap_build = -1;
    L.out = c->ms_out;
    L.nsets = 1;
    static const bool prev0 = getenv("SLK_MS_DEBUG_PREV0") != 0;  // timing aid for the debug launches only: results are invalid
    L.prev_n = prev0 ? 0 : -1;   // predecessor unknown: the step kernel waits for it before it visits anybody
    L.seed = c->seed; L.chain = c->chain_id;
    return L;
}

// Launch with programmatic stream serialization (see ms_launch_dependents / ms_wait_for_predecessor in
// slk_msampler.cuh): the kernel may become resident before its predecessor on the stream has finished.  Only for
// kernels that call ms_wait_for_predecessor() before touching anything the predecessor writes.
static void ms_launch_overlapped(const void* kernel, int grid, int threads, size_t smem, cudaStream_t stream, const SlkMsLaunch& L) {
    static const bool off = getenv("SLK_NO_PDL") != 0;               // tuning aid
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned) grid); cfg.blockDim = dim3((unsigned) threads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = off ? 0 : 1;
    void* args[1] = { (void*) &L };
    cudaLaunchKernelExC(&cfg, kernel, args);
}

// the incremental kernel of a sweep's steps (te and the stale masks are valid: a reset has run)
static void ms_launch_step(slk_chain* c, const SlkMsLaunch& L) {
    slk_plan* p = c->plan;
    const int grid = p->ms_grid * (L.nsets + (L.snap_build >= 0 ? 1 : 0));        // the snapshot set comes last
    if(2 * p->dev.F > 255) ms_launch_overlapped((const void*) slk_ms_step_kernel<true>, grid, 32, p->ms_step_smem, c->stream, L);
    else ms_launch_overlapped((const void*) slk_ms_step_kernel<false>, grid, 32, p->ms_step_smem, c->stream, L);
}

// The chain kernel runs while the next pair's likelihood kernel is already walking (see slk_ms_step_kernel) and has to
// become resident next to the running likelihood launch: it is what releases the next one.  SLK_MS_CHAIN_EXCLUSIVE=1
// (tuning aid, meaningful with SLK_MS_RUN_AHEAD=1 only) makes it ask for a whole SM's shared memory instead, so that no
// likelihood CTA is placed next to its eight 640-thread CTAs.
static void ms_launch_chain(slk_chain* c, const SlkMsLaunch& L) {
    static const char* force = getenv("SLK_MS_CHAIN_EXCLUSIVE");
    const bool exclusive = force && force[0] == '1';
    const size_t smem = exclusive ? (size_t)(slk::kSmemPerBlockMax - 4096) : (size_t) c->plan->ms_chain_smem;
    ms_launch_overlapped((const void*) slk_ms_chain_kernel, SLK_MS_CLUSTER, SLK_MS_CHAIN_THREADS, smem, c->stream, L);
}

static void ms_launch_likelihood(slk_chain* c, const SlkMsLaunch& L) {
    slk_plan* p = c->plan;
    const int grid = p->ms_grid * L.nsets;
    if(2 * p->dev.F > 255) slk_ms_likelihood_kernel<true><<<grid, 32, p->ms_smem, c->stream>>>(L);
    else slk_ms_likelihood_kernel<false><<<grid, 32, p->ms_smem, c->stream>>>(L);
}

// one or two consecutive steps of a sweep: the second meiosis's likelihood is evaluated under both
// outcomes of the first, so the pair costs one likelihood launch and one chain launch

// (q0, q1): the meioses of the chain kernel that precedes this launch on the stream (the previous pair of the
// sweep; q1 < 0: a single step), q0 < 0 when the predecessor is anything else
// run_ahead 2: (r0, r1) are the meioses of the pair before that one (< 0: none): the chain kernel of (q0, q1) and the
// likelihood launch of (q0, q1), which refreshes the label pairs (r0, r1) invalidated, may both still be running
// Forest snapshots of a launch (slk_ms_step_kernel): `use` / `build` = buffer (-1: none), the masks = te slots of the typed
// people the snapshot does NOT hold (anybody one of the window's meioses can affect)
struct MsSnapPlan {
    int use, build;
    const std::vector<uint32_t>* use_mask;
    const std::vector<uint32_t>* build_mask;
    MsSnapPlan() : use(-1), build(-1), use_mask(0), build_mask(0) {}
};

static int ms_steps(slk_chain* c, uint64_t iteration, int m0, int m1, int q0 = -1, int q1 = -1, int run_ahead = 1, int r0 = -1, int r1 = -1,
                    const MsSnapPlan& snap = MsSnapPlan()) {
    const SlkDevPlan& d = c->plan->dev;
    SlkMsLaunch L = ms_launch(c);
    L.iteration = iteration;
    L.prev_n = -1;
    L.run_ahead = run_ahead;
    if(q0 >= 0) {
        const int prev[4] = { q0, q1, run_ahead == 2 ? r0 : -1, run_ahead == 2 ? r1 : -1 };
        L.prev_n = 0;
        for(int j = 0; j < 4; ++j)
            if(prev[j] >= 0) { L.prev_person[L.prev_n] = d.F + prev[j] / 2; L.prev_parent[L.prev_n] = prev[j] % 2; ++L.prev_n; }
    }
    const int p0 = d.F + m0 / 2, a0 = m0 % 2;
    L.set_n[0] = 1; L.set_person[0][0] = p0; L.set_parent[0][0] = a0;
    L.nsteps = 1;
    L.step_person[0] = p0; L.step_parent[0] = a0; L.step_slot[0] = SLK_SLOT_MEIOSIS + (uint32_t) m0;
    if(m1 >= 0) {
        const int p1 = d.F + m1 / 2, a1 = m1 % 2;
        L.nsets = 3; L.nsteps = 2;
        L.set_n[1] = 1; L.set_person[1][0] = p1; L.set_parent[1][0] = a1;
        L.set_n[2] = 2; L.set_person[2][0] = p0; L.set_parent[2][0] = a0; L.set_person[2][1] = p1; L.set_parent[2][1] = a1;
        L.step_person[1] = p1; L.step_parent[1] = a1; L.step_slot[1] = SLK_SLOT_MEIOSIS + (uint32_t) m1;
        // the te refresh goes to the single-flip hypothesis with fewer slots below its person
        const std::vector<uint32_t>& dm = c->plan->host.ms_desc_mask;
        const int W = c->plan->host.ms_W;
        int n0 = 0, n1 = 0;
        for(int w = 0; w < W; ++w) { n0 += __builtin_popcount(dm[(size_t)(p0 - d.F) * W + w]); n1 += __builtin_popcount(dm[(size_t)(p1 - d.F) * W + w]); }
        L.refresh_set = n1 < n0 ? 1 : 0;
    }
    // the launch's record (visiting order, phase boundary, hypothesis slot masks), when it fits the parameter block
    {
        const slk::HostPlan& hp = c->plan->host;
        const int nt = d.ms.n_typed, W = hp.ms_W, words = slk_ms_pair_rec_words(nt, W);
        static const bool no_rec = getenv("SLK_MS_NO_REC") != 0;          // tuning aid: let the kernel derive it
        if(words <= SLK_MS_REC_WORDS && nt > 0 && !no_rec) {
            std::vector<uint32_t> prev((size_t) W, L.prev_n < 0 ? 0xffffffffu : 0u);
            auto add = [&](std::vector<uint32_t>& m, int person, int parent) {
                for(int w = 0; w < W; ++w) m[w] |= hp.ms_desc_mask[(size_t)(person - d.F) * W + w];
                const int kk = hp.ms_typed_index[person];
                if(kk >= 0) m[(2 * kk + parent) >> 5] |= 1u << ((2 * kk + parent) & 31);
            };
            for(int j = 0; j < L.prev_n; ++j) add(prev, L.prev_person[j], L.prev_parent[j]);
            uint32_t* rec = L.rec;
            memset(rec, 0, sizeof(uint32_t) * (size_t) words);
            for(int s = 0; s < L.nsets; ++s) {
                std::vector<uint32_t> hm((size_t) W, 0u);
                for(int j = 0; j < L.set_n[s]; ++j) add(hm, L.set_person[s][j], L.set_parent[s][j]);
                for(int w = 0; w < W; ++w) rec[3 + s * W + w] = hm[w];
            }
            auto touched = [&](const std::vector<uint32_t>& m, int k) { return ((m[k >> 4] >> (2 * (k & 15))) & 3u) != 0; };
            // the hypothesis sets visit everybody the snapshot they start from does not hold: first the people the
            // running chain kernels cannot affect, then the others
            const bool use = L.prev_n >= 0 && snap.use >= 0 && snap.use_mask;
            uint16_t* ord = (uint16_t*)(rec + 3 + 3 * W);
            int n = 0;
            for(int pass = 0; pass < 2; ++pass) {
                for(int k = 0; k < nt; ++k) {
                    if(use && !touched(*snap.use_mask, k)) continue;          // held by the snapshot
                    if(touched(prev, k) == (pass == 1)) ord[n++] = (uint16_t)(k | (hp.ms_typed_auto[k] ? 0x8000 : 0));
                }
                if(pass == 0) rec[0] = (uint32_t) n;
            }
            rec[1] = (uint32_t) n;
            if(use) L.snap_use = snap.use;
            // the snapshot set visits the people its snapshot holds
            if(snap.build >= 0 && snap.build_mask) {
                uint16_t* bord = (uint16_t*)(rec + 3 + 3 * W + (nt + 1) / 2);
                int nb = 0;
                for(int k = 0; k < nt; ++k)
                    if(!touched(*snap.build_mask, k)) bord[nb++] = (uint16_t)(k | (hp.ms_typed_auto[k] ? 0x8000 : 0));
                rec[2] = (uint32_t) nb;
                L.snap_build = snap.build;
            }
            L.rec_n = words;
        }
    }
    L.timeline = c->ms_timeline; L.tl_slot = c->ms_tl_slot;
    L.tl_cta_off = (c->ms_timeline && c->ms_tl_slot == c->ms_tl_cta_slot) ? 8 * ((int) c->plan->host.ms_ordering.size() + 2) : 0;
    static const bool full_kernel = getenv("SLK_MS_FULL_KERNEL") != 0;
    if(full_kernel) ms_launch_likelihood(c, L);                        // tuning aid: recompute every label each step
    else ms_launch_step(c, L);
    CU(cudaGetLastError());
    L.tl_slot = c->ms_tl_slot + 1;
    ms_launch_chain(c, L);
    CU(cudaGetLastError());
    return SLK_OK;
}

int slk_msampler_ordering(const slk_plan* plan, int32_t* out, int cap) {
    if(!plan) return 0;
    const std::vector<int32_t>& o = plan->host.ms_ordering;
    for(int i = 0; i < (int) o.size() && i < cap; ++i) out[i] = o[i];
    return (int) o.size();
}

int slk_msampler_reset(slk_chain* c) {
    int rc = ms_ready(c);
    if(rc != SLK_OK) return rc;
    SlkMsLaunch L = ms_launch(c);
    L.nsets = 1; L.set_n[0] = 0;
    ms_launch_likelihood(c, L);
    CU(cudaGetLastError());
    c->ms_cur_valid = true;
    return SLK_OK;
}

int slk_msampler_step(slk_chain* c, uint64_t iteration, int meiosis) {
    int rc = ms_ready(c);
    if(rc != SLK_OK) return rc;
    const SlkDevPlan& d = c->plan->dev;
    if(meiosis < 0 || meiosis >= 2 * (d.N - d.F)) return fail(SLK_ERR_INVALID, "meiosis %d out of range", meiosis);
    if(!c->ms_cur_valid) {
        rc = slk_msampler_reset(c);
        if(rc != SLK_OK) return rc;
    }
    return ms_steps(c, iteration, meiosis, -1);
}

int slk_msampler_sweep(slk_chain* c, uint64_t iteration) {
    int rc = ms_ready(c);
    if(rc != SLK_OK) return rc;
    std::vector<int32_t> order = c->plan->host.ms_ordering;
    const int n = (int) order.size();
    if(n == 0) return SLK_OK;
    // Fisher-Yates on the chain's Philox stream (the oracle's orc_ms_shuffle)
    for(int i = n - 1; i > 0; --i) {
        int j = (int)(slk_uniform(c->seed, c->chain_id, iteration, (uint32_t) i, SLK_SLOT_MSHUFFLE) * (double)(i + 1));
        if(j > i) j = i;
        std::swap(order[i], order[j]);
    }
    // MeiosisSampler::reset at the start of every sweep (markov_chain.cc:345)
    rc = slk_msampler_reset(c);
    // The likelihood walk of a pair runs ahead of the previous pair's chain kernel for every typed person that
    // pair cannot affect (slk_ms_step_kernel).  The first pair follows the reset kernel, which writes te.
    static const bool no_overlap = getenv("SLK_MS_NO_PREFIX") != 0;      // tuning aid
    static const int run_ahead = getenv("SLK_MS_RUN_AHEAD") ? atoi(getenv("SLK_MS_RUN_AHEAD")) : 2;   // tuning aid: 1 = one launch in flight
    // Forest snapshots (slk_ms_step_kernel): every K-th launch (a "leader", pair s) carries an extra set of CTAs that
    // walks the typed people no meiosis of pairs s-2 .. s+K+1 can affect and saves the forest; launches s+2 .. s+K+1
    // -- which start after the leader has completed -- begin their walks from it and visit the other people only.
    // Two buffers alternate: the launches that read one have completed before the leader after next starts.
    static const int snap_k = getenv("SLK_MS_SNAPSHOT") ? atoi(getenv("SLK_MS_SNAPSHOT")) : 4;    // tuning aid: 0 = off
    const int np = (n + 1) / 2;
    std::vector<MsSnapPlan> snaps((size_t) np);
    std::vector<std::vector<uint32_t> > masks;
    static const bool snap_forced = getenv("SLK_MS_SNAPSHOT") != 0;
    // (a short walk is not worth saving, unless the tuning aid asks for it: the tests do, on the small pedigrees)
    if(snap_k > 0 && run_ahead == 2 && !no_overlap && (snap_forced || c->plan->dev.ms.n_typed >= 32)) {
        const slk::HostPlan& hp = c->plan->host;
        const int W = hp.ms_W, nt = c->plan->dev.ms.n_typed, F = c->plan->dev.F;
        masks.reserve((size_t) np / snap_k + 1);
        for(int s = 0; s + 2 < np; s += snap_k) {
            std::vector<uint32_t> m((size_t) W, 0u);
            for(int j = std::max(0, s - 2); j <= std::min(np - 1, s + snap_k + 1); ++j)
                for(int h = 0; h < 2 && 2 * j + h < n; ++h) {
                    const int person = F + order[2 * j + h] / 2, parent = order[2 * j + h] % 2;
                    for(int w = 0; w < W; ++w) m[w] |= hp.ms_desc_mask[(size_t)(person - F) * W + w];
                    const int kk = hp.ms_typed_index[person];
                    if(kk >= 0) m[(2 * kk + parent) >> 5] |= 1u << ((2 * kk + parent) & 31);
                }
            int held = 0;
            for(int k = 0; k < nt; ++k) if(((m[k >> 4] >> (2 * (k & 15))) & 3u) == 0) ++held;
            if(held < (snap_forced ? 1 : 8)) continue;               // not worth a snapshot
            masks.push_back(m);
            const int buf = (s / snap_k) & 1;
            snaps[s].build = buf; snaps[s].build_mask = &masks.back();
            for(int j = s + 2; j <= std::min(np - 1, s + snap_k + 1); ++j) { snaps[j].use = buf; snaps[j].use_mask = &masks.back(); }
        }
    }
    int q0 = -1, q1 = -1, r0 = -1, r1 = -1;
    for(int j = 0; j < n && rc == SLK_OK; j += 2) {
        const int m0 = order[j], m1 = j + 1 < n ? order[j + 1] : -1;
        c->ms_tl_slot = j;
        rc = ms_steps(c, iteration, m0, m1, no_overlap ? -1 : q0, q1, (no_overlap || run_ahead != 2) ? 1 : 2, r0, r1, snaps[j / 2]);
        r0 = q0; r1 = q1;
        q0 = m0; q1 = m1;
    }
    return rc;
}

// Tuning aid: one M-sweep with %globaltimer stamps.  out[8 j + 0..4] = likelihood launch of the pair that starts at
// meiosis j of the order (start, walk, before wait, after wait, end; CTA 0), out[8 (j + 1) + 0..3] = its chain launch
// (start, before wait, after wait, end); from out[8 (n + 2)] on, (start, end << 10 | SM) of every CTA of the likelihood
// launch at order position `cta_pair`.  cap >= 8 (n + 2) + 8 * ceil(M / 32) words (n = meioses of a sweep).
int slk_debug_msampler_timeline(slk_chain* c, uint64_t iteration, int cta_pair, unsigned long long* out, int cap) {
    int rc = ms_ready(c);
    if(rc != SLK_OK) return rc;
    const int n = (int) c->plan->host.ms_ordering.size();
    const int words = 8 * (n + 2) + 2 * 4 * c->plan->ms_grid;          // up to four sets of CTAs (the fourth: a snapshot set)
    if(!out || cap < words) return fail(SLK_ERR_INVALID, "slk_debug_msampler_timeline: %d words needed", words);
    unsigned long long* tl = 0;
    CU(cudaMalloc((void**) &tl, sizeof(unsigned long long) * words));
    CU(cudaMemsetAsync(tl, 0, sizeof(unsigned long long) * words, c->stream));
    c->ms_timeline = tl; c->ms_tl_cta_slot = cta_pair;
    rc = slk_msampler_sweep(c, iteration);
    c->ms_timeline = 0;
    cudaError_t e = cudaMemcpyAsync(out, tl, sizeof(unsigned long long) * words, cudaMemcpyDeviceToHost, c->stream);
    if(e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(tl);
    if(e != cudaSuccess) return fail(SLK_ERR_CUDA, "timeline: %s", cudaGetErrorString(e));
    return rc;
}

int slk_dg_likelihood(slk_chain* c, double* log_likelihood) {
    if(!log_likelihood) return fail(SLK_ERR_INVALID, "null argument");
    int rc = ms_ready(c);
    if(rc != SLK_OK) return rc;
    if(!c->ms_cur_valid) {
        rc = slk_msampler_reset(c);
        if(rc != SLK_OK) return rc;
    }
    SlkMsLaunch L = ms_launch(c);
    slk_ms_dg_likelihood_kernel<<<1, 1024, 0, c->stream>>>(L);
    CU(cudaGetLastError());
    double h[2] = {0.0, 0.0};
    CU(cudaMemcpyAsync(h, c->ms_out, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    rc = check_device_error(c);
    if(rc != SLK_OK) return rc;
    // descent_graph.cc:150-169: log_product(marker_transmission + recombination, sum prior)
    *log_likelihood = (h[0] == -DBL_MAX) ? -DBL_MAX : (c->plan->dev.marker_transmission + h[1]) + h[0];
    return SLK_OK;
}

int slk_sweep_is_lsampler(const slk_chain* c, uint64_t iteration, double lsampler_prob) {
    if(!c) return 1;
    return slk_uniform(c->seed, c->chain_id, iteration, 0u, SLK_SLOT_KIND) < lsampler_prob ? 1 : 0;
}

int slk_debug_fag(slk_chain* c, int meiosis, double* lnlik, int32_t* edges) {
    if(!lnlik) return fail(SLK_ERR_INVALID, "null argument");
    int rc = ms_ready(c);
    if(rc != SLK_OK) return rc;
    const SlkDevPlan& d = c->plan->dev;
    if(meiosis >= 2 * (d.N - d.F)) return fail(SLK_ERR_INVALID, "meiosis %d out of range", meiosis);
    SlkMsLaunch L = ms_launch(c);
    int32_t* dev_edges = 0;
    if(edges) {
        CU(cudaMalloc((void**) &dev_edges, sizeof(int32_t) * (size_t) d.M * 2 * d.N));
        L.dump_edges = dev_edges;
    }
    L.nsets = 1;
    if(meiosis >= 0) { L.set_n[0] = 1; L.set_person[0][0] = d.F + meiosis / 2; L.set_parent[0][0] = meiosis % 2; }
    else { L.set_n[0] = 0; c->ms_cur_valid = true; }         // the no-flip launch refreshes ms_lncur
    ms_launch_likelihood(c, L);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(lnlik, meiosis >= 0 ? c->ms_lnl : c->ms_lncur, sizeof(double) * d.M, cudaMemcpyDeviceToHost, c->stream));
    if(edges) CU(cudaMemcpyAsync(edges, dev_edges, sizeof(int32_t) * (size_t) d.M * 2 * d.N, cudaMemcpyDeviceToHost, c->stream));
    rc = check_device_error(c);
    cudaFree(dev_edges);
    return rc;
}

int slk_debug_msampler_trace(slk_chain* c, int meiosis0, int meiosis1, long long* stamps) {
    if(!stamps) return fail(SLK_ERR_INVALID, "null argument");
    int rc = ms_ready(c);
    if(rc != SLK_OK) return rc;
    const SlkDevPlan& d = c->plan->dev;
    long long* dev = 0;
    CU(cudaMalloc((void**) &dev, sizeof(long long) * 160));
    CU(cudaMemsetAsync(dev, 0, sizeof(long long) * 160, c->stream));
    SlkMsLaunch L = ms_launch(c);
    L.trace = dev;
    L.nsets = 3;
    const int p0 = d.F + meiosis0 / 2, a0 = meiosis0 % 2, p1 = d.F + meiosis1 / 2, a1 = meiosis1 % 2;
    L.set_n[0] = 1; L.set_person[0][0] = p0; L.set_parent[0][0] = a0;
    L.set_n[1] = 1; L.set_person[1][0] = p1; L.set_parent[1][0] = a1;
    L.set_n[2] = 2; L.set_person[2][0] = p0; L.set_parent[2][0] = a0; L.set_person[2][1] = p1; L.set_parent[2][1] = a1;
    L.nsteps = 2;
    L.step_person[0] = p0; L.step_parent[0] = a0; L.step_slot[0] = SLK_SLOT_MEIOSIS + (uint32_t) meiosis0;
    L.step_person[1] = p1; L.step_parent[1] = a1; L.step_slot[1] = SLK_SLOT_MEIOSIS + (uint32_t) meiosis1;
    if(!c->ms_cur_valid) {
        rc = slk_msampler_reset(c);
        if(rc != SLK_OK) return rc;
    }
    ms_launch_step(c, L);
    CU(cudaGetLastError());
    if(meiosis0 != meiosis1) {
        // the chain kernel's stamps go to rows 12.. (CTA 0) and 15.. (last CTA) of the same buffer
        SlkMsLaunch L2 = L;
        L2.trace = dev + 96;
        ms_launch_chain(c, L2);
        CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(stamps, dev, sizeof(long long) * 160, cudaMemcpyDeviceToHost, c->stream));
    rc = check_device_error(c);
    cudaFree(dev);
    return rc;
}

int slk_debug_msampler_launch(slk_chain* c, int meiosis0, int meiosis1, int which, int reps) {
    int rc = ms_ready(c);
    if(rc != SLK_OK) return rc;
    const SlkDevPlan& d = c->plan->dev;
    if(meiosis0 < 0 || meiosis1 < 0 || meiosis0 >= 2 * (d.N - d.F) || meiosis1 >= 2 * (d.N - d.F)) return fail(SLK_ERR_INVALID, "meiosis out of range");
    if(!c->ms_cur_valid) {
        rc = slk_msampler_reset(c);
        if(rc != SLK_OK) return rc;
    }
    SlkMsLaunch L = ms_launch(c);
    const int p0 = d.F + meiosis0 / 2, a0 = meiosis0 % 2, p1 = d.F + meiosis1 / 2, a1 = meiosis1 % 2;
    L.nsets = 3; L.nsteps = 2;
    L.set_n[0] = 1; L.set_person[0][0] = p0; L.set_parent[0][0] = a0;
    L.set_n[1] = 1; L.set_person[1][0] = p1; L.set_parent[1][0] = a1;
    L.set_n[2] = 2; L.set_person[2][0] = p0; L.set_parent[2][0] = a0; L.set_person[2][1] = p1; L.set_parent[2][1] = a1;
    L.step_person[0] = p0; L.step_parent[0] = a0; L.step_slot[0] = SLK_SLOT_MEIOSIS + (uint32_t) meiosis0;
    L.step_person[1] = p1; L.step_parent[1] = a1; L.step_slot[1] = SLK_SLOT_MEIOSIS + (uint32_t) meiosis1;
    if(which == 0) {
        for(int r = 0; r < reps; ++r) ms_launch_step(c, L);
    }
    else if(which == 2) {
        for(int r = 0; r < reps; ++r) ms_launch_likelihood(c, L);      // the full (non-incremental) kernel
    }
    else {
        ms_launch_step(c, L);                          // the chain kernel needs this pair's likelihoods
        for(int r = 0; r < reps; ++r) {
            L.iteration = (uint64_t) r;
            ms_launch_chain(c, L);
        }
        c->ms_cur_valid = false;                       // repeated sampling of one pair: the graph is still legal, ln cur is stale
    }
    CU(cudaGetLastError());
    return SLK_OK;
}

int slk_debug_msampler_state(slk_chain* c, double* fb, double* lncur) {
    int rc = ms_ready(c);
    if(rc != SLK_OK) return rc;
    const SlkDevPlan& d = c->plan->dev;
    if(fb)    CU(cudaMemcpyAsync(fb, c->ms_fb, sizeof(double) * 2 * d.M, cudaMemcpyDeviceToHost, c->stream));
    if(lncur) CU(cudaMemcpyAsync(lncur, c->ms_lncur, sizeof(double) * d.M, cudaMemcpyDeviceToHost, c->stream));
    return check_device_error(c);
}

}  // extern "C"

// ---- ELOD --------------------------------------------------------------------------------

namespace {

// log-sum-exp of prob[0..n): one CTA, fixed-order tree (logarithms.cc:14-23 applied pairwise)
__global__ void __launch_bounds__(1024) slk_logsumexp_kernel(const double* prob, long long n, double* out) {
    __shared__ double s_m[1024], s_s[1024];
    __shared__ long long s_c[1024];
    const int t = threadIdx.x, T = blockDim.x;
    double m = -DBL_MAX; long long cnt = 0;
    for(long long i = t; i < n; i += T) { const double v = prob[i]; if(v > -DBL_MAX) { if(v > m) m = v; ++cnt; } }
    double sum = 0.0;
    for(long long i = t; i < n; i += T) { const double v = prob[i]; if(v > -DBL_MAX) sum += exp(v - m); }
    s_m[t] = m; s_s[t] = sum; s_c[t] = cnt;
    __syncthreads();
    for(int d = T >> 1; d > 0; d >>= 1) {
        if(t < d) {
            const double m1 = s_m[t], m2 = s_m[t + d];
            const double mm = m1 > m2 ? m1 : m2;
            double acc = 0.0;
            if(s_c[t]) acc += s_s[t] * exp(m1 - mm);
            if(s_c[t + d]) acc += s_s[t + d] * exp(m2 - mm);
            s_m[t] = mm; s_s[t] = acc; s_c[t] += s_c[t + d];
        }
        __syncthreads();
    }
    if(t == 0) { out[0] = s_c[0] ? s_m[0] + log(s_s[0]) : -DBL_MAX; out[1] = (double) s_c[0]; }
}

struct ElodScratch {
    uint8_t* dg; double* gs; double* prob; double* red; int* err;
    ElodScratch() : dg(0), gs(0), prob(0), red(0), err(0) {}
    ~ElodScratch() { cudaFree(dg); cudaFree(gs); cudaFree(prob); cudaFree(red); cudaFree(err); }
};

// the three start_from launches of a chunk of replicates on the wide graph (locus_sampler2.cc:214-241
// with starting_locus = 1): the middle locus alone, then the left marker given its right neighbour, then
// the right marker given its left neighbour
int elod_sample_chunk(slk_plan* sp, const ElodScratch& sc, uint64_t seed, uint32_t chain_id, long long first, int n, cudaStream_t st) {
    const SlkDevPlan& d = sp->dev;
    const int tpc = d.ls.cta_threads / d.ls.team_threads;
    for(int pass = 0; pass < 3; ++pass) {
        SlkLaunch L;
        memset(&L, 0, sizeof(L));
        L.plan = d; L.dgp = sc.dg; L.gscratch = sc.gs; L.err = sc.err;
        L.seed = seed; L.chain = chain_id; L.iteration = 0;
        L.dump_k = -1; L.si_start = -1;
        L.period = 3; L.window = 3; L.nunits = n;
        // Philox is keyed by the row of the wide graph: offset the rows by the first replicate so that a
        // replicate's draws do not depend on the chunking
        L.offset = (pass == 0) ? 1 : (pass == 1 ? 0 : 2);
        L.ignore_left = (pass == 2) ? 0 : 1;
        L.ignore_right = (pass == 1) ? 0 : 1;
        L.row_base = first * 3;
        launch_ls_by_team(d.ls.team_threads, L, grid_for(n, tpc, sp->ls_blocks_per_sm, sp->sm_count), d.ls.cta_threads, sp->ls_cta_smem, st);
        CU(cudaGetLastError());
    }
    return SLK_OK;
}

const int kElodChunk = 1 << 16;

int elod_alloc(slk_plan* sp, slk_plan* tp, ElodScratch& sc, int chunk) {
    const SlkDevPlan& d = sp->dev;
    const SlkDevPlan& t = tp->dev;
    size_t ls_teams = (size_t) sp->ls_blocks_per_sm * sp->sm_count * (d.ls.cta_threads / d.ls.team_threads);
    size_t lod_teams = (size_t) tp->lod_blocks_per_sm * tp->sm_count * (t.lod.cta_threads / t.lod.team_threads);
    size_t gs = std::max(ls_teams * (size_t)(d.ls.arena_doubles - d.ls.smem_doubles),
                         lod_teams * (size_t)(t.lod.arena_doubles - t.lod.smem_doubles));
    CU(cudaMalloc((void**) &sc.dg, (size_t) chunk * 3 * d.N));
    CU(cudaMemset(sc.dg, 0, (size_t) chunk * 3 * d.N));
    CU(cudaMalloc((void**) &sc.gs, sizeof(double) * std::max<size_t>(gs, 2)));
    CU(cudaMalloc((void**) &sc.prob, sizeof(double) * chunk));
    CU(cudaMalloc((void**) &sc.red, sizeof(double) * 2));
    CU(cudaMalloc((void**) &sc.err, sizeof(int) * 4));
    CU(cudaMemset(sc.err, 0, sizeof(int) * 4));
    return SLK_OK;
}

int elod_check(slk_plan* sp, slk_plan* tp) {
    if(!sp || (tp && (sp->device != tp->device || sp->dev.N != tp->dev.N || sp->dev.F != tp->dev.F)))
        return fail(SLK_ERR_INVALID, "slk_elod: the two plans must describe one pedigree on one device");
    if(sp->dev.M != 3 || (tp && (tp->dev.M != 2 || tp->dev.nlod != 1)))
        return fail(SLK_ERR_INVALID, "slk_elod: sampler plan must have 3 loci, trait plan 2 loci and n_lod = 1");
    return SLK_OK;
}

}  // namespace

extern "C" int slk_elod_run(slk_plan* sp, slk_plan* tp, uint64_t seed, uint32_t chain_id, int64_t replicates,
                            double* log_sum, int64_t* count, double* prob_out) {
    if(!tp || !log_sum || !count || replicates < 1) return fail(SLK_ERR_INVALID, "bad argument");
    int rc = elod_check(sp, tp);
    if(rc != SLK_OK) return rc;
    CU(cudaSetDevice(sp->device));
    const int chunk = (int) std::min<int64_t>(replicates, kElodChunk);
    ElodScratch sc;
    rc = elod_alloc(sp, tp, sc, chunk);
    if(rc != SLK_OK) return rc;
    cudaStream_t st = 0;
    const SlkDevPlan& t = tp->dev;
    double total = -DBL_MAX;
    long long added = 0;
    for(int64_t first = 0; first < replicates; first += chunk) {
        const int n = (int) std::min<int64_t>(chunk, replicates - first);
        rc = elod_sample_chunk(sp, sc, seed, chain_id, first, n, st);
        if(rc != SLK_OK) return rc;
        // Peeler::process of every replicate's two marker rows (elod.cc:58-61)
        SlkLaunch L;
        memset(&L, 0, sizeof(L));
        L.plan = t; L.dgp = sc.dg; L.gscratch = sc.gs; L.err = sc.err;
        L.dump_k = -1; L.si_start = -1; L.window = 2;
        L.period = 3; L.lod_row0 = 0; L.lod_row1 = 2;
        L.accumulate = 0; L.nunits = n; L.dump_prob = sc.prob; L.dump_result = sc.prob;   // result is overwritten by prob
        const int tpc = t.lod.cta_threads / t.lod.team_threads;
        launch_lod_by_team(t.lod.team_threads, L, grid_for(n, tpc, tp->lod_blocks_per_sm, tp->sm_count), t.lod.cta_threads, tp->lod_cta_smem, st);
        CU(cudaGetLastError());
        slk_logsumexp_kernel<<<1, 1024, 0, st>>>(sc.prob, n, sc.red);
        CU(cudaGetLastError());
        double h[2]; int herr[2];
        CU(cudaMemcpyAsync(h, sc.red, sizeof(h), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(herr, sc.err, sizeof(herr), cudaMemcpyDeviceToHost, st));
        if(prob_out) CU(cudaMemcpyAsync(prob_out + first, sc.prob, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if(herr[0] != 0) {
            if(herr[0] == SLK_ERR_ZERO_LIKELIHOOD) return fail(herr[0], "likelihood is zero at locus %d (check penetrance function?)", herr[1] % 3);
            return fail(herr[0], "intermediate state had a likelihood of 0.0 or less (replicate %lld)", (long long) first + herr[1]);
        }
        // LODscores::add across chunks (lod_score.h:74-80)
        if(h[1] > 0) {
            total = (total == -DBL_MAX) ? h[0] : log(exp(total - h[0]) + 1.0) + h[0];
            added += (long long) h[1];
        }
    }
    *log_sum = total;
    *count = added;
    return SLK_OK;
}

extern "C" int slk_debug_elod_graphs(slk_plan* sp, uint64_t seed, uint32_t chain_id, int64_t first, int32_t n, int32_t* dg) {
    if(!dg || n < 1 || first < 0) return fail(SLK_ERR_INVALID, "bad argument");
    int rc = elod_check(sp, 0);
    if(rc != SLK_OK) return rc;
    CU(cudaSetDevice(sp->device));
    ElodScratch sc;
    rc = elod_alloc(sp, sp, sc, n);
    if(rc != SLK_OK) return rc;
    rc = elod_sample_chunk(sp, sc, seed, chain_id, first, n, 0);
    if(rc != SLK_OK) return rc;
    const size_t cells = (size_t) n * 3 * sp->dev.N;
    int32_t* staging = 0;
    CU(cudaMalloc((void**) &staging, cells * 2 * sizeof(int32_t)));
    slk_dg_unpack_kernel<<<(unsigned)((cells + 255) / 256), 256>>>(sc.dg, staging, cells);
    CU(cudaGetLastError());
    CU(cudaMemcpy(dg, staging, cells * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost));
    cudaFree(staging);
    return SLK_OK;
}

extern "C" {

// ---- parity hooks ----------------------------------------------------------------------

static int debug_ls(slk_chain* c, uint64_t iteration, int locus, int il, int ir, bool forward_only,
                    double* matrices, double* presums, int32_t* pmk, double* dist4, double* result) {
    if(!c) return fail(SLK_ERR_INVALID, "null chain");
    slk_plan* p = c->plan;
    const SlkDevPlan& d = p->dev;
    const slk::HostPlan& hp = p->host;
    if(locus < 0 || locus >= d.M) return fail(SLK_ERR_INVALID, "locus out of range");
    CU(cudaSetDevice(p->device));
    int rc = ensure_debug(c);
    if(rc != SLK_OK) return rc;
    uint8_t* saved = 0;
    if(forward_only) {
        // the kernel always completes the update; keep the row so the call has no side effect
        CU(cudaMalloc((void**) &saved, d.N));
        CU(cudaMemcpyAsync(saved, c->dgp + (size_t) locus * d.N, d.N, cudaMemcpyDeviceToDevice, c->stream));
    }
    CU(cudaMemsetAsync(c->dbg_mat, 0, sizeof(double) * hp.sum_cells, c->stream));
    CU(cudaMemsetAsync(c->dbg_pre, 0, sizeof(double) * hp.sum_presum, c->stream));
    CU(cudaMemsetAsync(c->dbg_dist4, 0, sizeof(double) * 4 * hp.nops, c->stream));
    CU(cudaMemsetAsync(c->dbg_pmk, 0xff, sizeof(int) * hp.N, c->stream));
    SlkLaunch L = base_launch(c);
    L.iteration = iteration;
    L.window = d.M + 1;
    L.offset = locus;
    L.nunits = 1;
    L.ignore_left = il; L.ignore_right = ir;
    L.dump_mat = c->dbg_mat; L.dump_pre = c->dbg_pre; L.dump_dist4 = c->dbg_dist4; L.dump_pmk = c->dbg_pmk;
    L.dump_result = c->dbg_res;
    launch_ls_debug(L, d.ls.prog_smem_bytes + d.ls.team_smem_bytes, c->stream);
    CU(cudaGetLastError());
    if(forward_only) {
        CU(cudaMemcpyAsync(c->dgp + (size_t) locus * d.N, saved, d.N, cudaMemcpyDeviceToDevice, c->stream));
    }
    if(matrices) CU(cudaMemcpyAsync(matrices, c->dbg_mat, sizeof(double) * hp.sum_cells, cudaMemcpyDeviceToHost, c->stream));
    if(presums)  CU(cudaMemcpyAsync(presums, c->dbg_pre, sizeof(double) * hp.sum_presum, cudaMemcpyDeviceToHost, c->stream));
    if(pmk)      CU(cudaMemcpyAsync(pmk, c->dbg_pmk, sizeof(int) * hp.N, cudaMemcpyDeviceToHost, c->stream));
    if(dist4)    CU(cudaMemcpyAsync(dist4, c->dbg_dist4, sizeof(double) * 4 * hp.nops, cudaMemcpyDeviceToHost, c->stream));
    if(result)   CU(cudaMemcpyAsync(result, c->dbg_res, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if(saved) cudaFree(saved);
    if(forward_only) {
        // a zero likelihood is a legitimate answer for the forward hook
        cudaMemsetAsync(c->err, 0, 4 * sizeof(int), c->stream);
        return SLK_OK;
    }
    return check_device_error(c);
}

int slk_debug_lsampler_forward(slk_chain* c, int locus, int ignore_left, int ignore_right,
                               double* matrices, double* presums, double* result) {
    return debug_ls(c, 0, locus, ignore_left, ignore_right, true, matrices, presums, 0, 0, result);
}

int slk_debug_lsampler_step(slk_chain* c, uint64_t iteration, int locus, int ignore_left, int ignore_right,
                            int32_t* pmk, double* dist4, double* result) {
    return debug_ls(c, iteration, locus, ignore_left, ignore_right, false, 0, 0, pmk, dist4, result);
}

int slk_debug_lod_interval(slk_chain* c, int interval, double* result, double* prob, int dump_k, double* matrices) {
    if(!c || !result || !prob) return fail(SLK_ERR_INVALID, "null argument");
    slk_plan* p = c->plan;
    const SlkDevPlan& d = p->dev;
    const slk::HostPlan& hp = p->host;
    if(interval < 0 || interval >= d.M - 1) return fail(SLK_ERR_INVALID, "interval out of range");
    CU(cudaSetDevice(p->device));
    int rc = ensure_debug(c);
    if(rc != SLK_OK) return rc;
    const int tpc = d.lod.cta_threads / d.lod.team_threads;
    SlkLaunch L = base_launch(c);
    L.accumulate = 0;
    L.unit_base = interval * d.nlod;
    L.nunits = d.nlod;
    L.dump_result = c->dbg_res;
    L.dump_prob = c->dbg_prob;
    (void) tpc;
    launch_lod_debug(L, L.nunits, d.lod.prog_smem_bytes + d.lod.team_smem_bytes, c->stream);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(result, c->dbg_res, sizeof(double) * d.nlod, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(prob, c->dbg_prob, sizeof(double) * d.nlod, cudaMemcpyDeviceToHost, c->stream));
    if(matrices && dump_k >= 0 && dump_k < d.nlod) {
        CU(cudaMemsetAsync(c->dbg_mat, 0, sizeof(double) * hp.sum_cells, c->stream));
        SlkLaunch L2 = L;
        L2.unit_base = interval * d.nlod + dump_k;
        L2.nunits = 1;
        L2.dump_mat = c->dbg_mat;
        L2.dump_result = c->dbg_res + d.nlod;
        L2.dump_prob = c->dbg_prob + d.nlod;
        launch_lod_debug(L2, 1, d.lod.prog_smem_bytes + d.lod.team_smem_bytes, c->stream);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(matrices, c->dbg_mat, sizeof(double) * hp.sum_cells, cudaMemcpyDeviceToHost, c->stream));
    }
    return check_device_error(c);
}

int slk_debug_lsampler_trace(slk_chain* c, uint64_t iteration, int offset, long long* stamps, int cap) {
    if(!c || !stamps || cap < 8) return fail(SLK_ERR_INVALID, "bad argument");
    slk_plan* p = c->plan;
    const SlkDevPlan& d = p->dev;
    CU(cudaSetDevice(p->device));
    const int n = d.ls.n_flevels + d.ls.n_blevels + 3;
    if(n > cap) return fail(SLK_ERR_INVALID, "need room for %d stamps", n);
    long long* dev = 0;
    CU(cudaMalloc((void**) &dev, sizeof(long long) * n));
    CU(cudaMemsetAsync(dev, 0, sizeof(long long) * n, c->stream));
    SlkLaunch L = base_launch(c);
    L.iteration = iteration;
    L.window = 2;
    L.offset = offset & 1;
    L.nunits = (d.M - L.offset + 1) / 2;
    L.trace = dev;
    const int tpc = d.ls.cta_threads / d.ls.team_threads;
    launch_ls_by_team(d.ls.team_threads, L, grid_for(L.nunits, tpc, p->ls_blocks_per_sm, p->sm_count), d.ls.cta_threads, p->ls_cta_smem, c->stream);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(stamps, dev, sizeof(long long) * n, cudaMemcpyDeviceToHost, c->stream));
    int rc = check_device_error(c);
    cudaFree(dev);
    return rc == SLK_OK ? n : rc;
}

int slk_debug_philox(int device, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    if(slk_device_count() <= device) return fail(SLK_ERR_NO_DEVICE, "no usable CUDA device %d", device);
    CU(cudaSetDevice(device));
    uint32_t* d = 0;
    CU(cudaMalloc((void**) &d, 16));
    slk_philox_kernel<<<1, 1>>>(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], d);
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, d, 16, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return SLK_OK;
}

int slk_debug_uniform(int device, uint64_t seed, uint32_t chain, uint64_t iteration, uint32_t locus,
                      uint32_t slot, double* out) {
    if(slk_device_count() <= device) return fail(SLK_ERR_NO_DEVICE, "no usable CUDA device %d", device);
    CU(cudaSetDevice(device));
    double* d = 0;
    CU(cudaMalloc((void**) &d, 8));
    slk_uniform_kernel<<<1, 1>>>(seed, chain, iteration, locus, slot, d);
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, d, 8, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return SLK_OK;
}

int slk_measure_fp64_peak(int device, double* tflops) {
    if(!tflops) return fail(SLK_ERR_INVALID, "null argument");
    if(slk_device_count() <= device) return fail(SLK_ERR_NO_DEVICE, "no usable CUDA device %d", device);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 4, threads = 512, iters = 1 << 14;
    double* d = 0;
    CU(cudaMalloc((void**) &d, sizeof(double) * blocks * threads));
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
    double best = 0.0;
    for(int rep = 0; rep < 6; ++rep) {
        CU(cudaEventRecord(a));
        slk_fp64_peak_kernel<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
        CU(cudaEventRecord(b));
        CU(cudaEventSynchronize(b));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, a, b));
        double fl = 2.0 * 8.0 * (double) iters * blocks * threads;
        double tf = fl / (ms * 1e-3) / 1e12;
        if(rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(d);
    *tflops = best;
    return SLK_OK;
}

}  // extern "C"
