// gpu.cc -- GPULodscores and GPUMarkovChain on top of the C ABI.  Same public surface as the
// reference's gpu_lodscores.h:63-120 and gpu_markov_chain.h:76,125; errors print and abort()
// as CUDA_CALLANDTEST does (gpu_lodscores.cc:30-37).  Unlike the reference's destructor
// (gpu_lodscores.h:95-99) nothing here calls cudaDeviceReset(): several instances can coexist
// in one process, one per GPU.
#include "swiftlink_host.h"
#include "slk_philox.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cfloat>
#include <cstring>

namespace swiftlink {

static void die_on(int rc, const char* what) {
    if(rc != SLK_OK) {
        fprintf(stderr, "error: %s: %s\n", what, slk_last_error());
        abort();
    }
}

GPULodscores::GPULodscores(Pedigree* ped, GeneticMap* map, PeelSequenceGenerator* psg, struct mcmc_options options, double trait_prob) :
    ped(ped), map(map), psg(psg), options(options), trait_likelihood(trait_prob), plan(0), chain(0), owns_plan(true) {
    flatten_problem(*ped, *map, *psg, options.sex_linked, flat);
    die_on(slk_plan_create(&flat.desc, options.device, &plan), "GPULodscores: plan");
    die_on(slk_chain_create(plan, options.seed, 0, &chain), "GPULodscores: chain");
}

GPULodscores::~GPULodscores() {
    slk_chain_destroy(chain);
    if(owns_plan) slk_plan_destroy(plan);
}

void GPULodscores::calculate(DescentGraph& dg) {
    // the upload is synchronous with respect to the caller's buffer, the scoring launch is not:
    // CPU sampling overlaps GPU scoring exactly as in the reference
    die_on(slk_dg_upload(chain, dg.get_internal_ptr()), "GPULodscores::calculate (upload)");
    die_on(slk_lodscore_accumulate(chain), "GPULodscores::calculate (launch)");
}

void GPULodscores::block_until_finished() {
    die_on(slk_chain_sync(chain), "GPULodscores::block_until_finished");
}

void GPULodscores::get_results(LODscores* lod) {
    std::vector<double> raw(lod->num_lodscores());
    int32_t count = 0;
    die_on(slk_lodscore_read(chain, raw.data(), &count), "GPULodscores::get_results");
    lod->set_trait_prob(trait_likelihood);
    lod->set_count((unsigned int) count);
    for(unsigned int i = 0; i < raw.size(); ++i) lod->set(i, raw[i]);
}

GPUMarkovChain::GPUMarkovChain(Pedigree* ped, GeneticMap* map, PeelSequenceGenerator* psg, struct mcmc_options options,
                               int sequence_num, double temperature) :
    ped(ped), map(map), psg(psg), options(options), heated(*map), temperature(temperature), plan(0), chain(0),
    seq_num(sequence_num), trait_prob(0.0), scoring_started(false), coda(0) {
    // MarkovChain::_init: "heat up the map" (markov_chain.cc:33-34).  Thetas and allele frequencies move
    // towards 0.5; the genotype priors cached per person stay cold, as in the reference (Snp::prob is
    // not recomputed by set_minor_freq, genetic_map.h:43-47)
    if(temperature != 1.0) heated.set_temperature(temperature);
    flatten_problem(*ped, heated, *psg, options.sex_linked, flat);
    die_on(slk_plan_create(&flat.desc, options.device, &plan), "GPUMarkovChain: plan");
    die_on(slk_chain_create(plan, options.seed, (uint32_t) sequence_num, &chain), "GPUMarkovChain: chain");
}

GPUMarkovChain::~GPUMarkovChain() {
    slk_chain_destroy(chain);
    slk_plan_destroy(plan);
}

double GPUMarkovChain::calc_trait_prob() {
    double v = 0.0;
    die_on(slk_trait_likelihood(plan, &v), "GPUMarkovChain::calc_trait_prob");
    return v;
}

double GPUMarkovChain::sequential_imputation(DescentGraph& dg, int iterations) {
    const int M = (int) map->num_markers();
    if(iterations == 0) {
        die_on(slk_lsampler_locus_by_locus(chain, SLK_SI_FIRST_RUN), "sequential imputation (locus by locus)");
        die_on(slk_dg_download(chain, dg.get_internal_ptr()), "sequential imputation (download)");
        return 0.0;
    }
    // the reference runs its `iterations` walks on the OpenMP threads and keeps the best (sequential_imputation.cc:
    // 47-115); here they are teams of one launch (slk_sequential_imputation_batch)
    HostRng rng(options.seed ^ (0x5349ull << 32) ^ (uint64_t) seq_num);
    std::vector<int32_t> starts((size_t) iterations);
    for(int run = 0; run < iterations; ++run) starts[run] = (int32_t) rng.uniform_int(M);
    std::vector<double> weights((size_t) iterations);
    int32_t best_run = -1;
    // run numbers live in their own part of the Philox counter space (slk_philox.cuh): the MCMC iterations that
    // follow never replay the draws that built the start graph, whatever the burn-in
    die_on(slk_sequential_imputation_batch(chain, SLK_SI_FIRST_RUN, iterations, starts.data(), weights.data(), &best_run), "sequential imputation");
    const double best = best_run >= 0 ? weights[best_run] : -1e300;
    die_on(slk_dg_download(chain, dg.get_internal_ptr()), "sequential imputation (download)");
    printf("starting likelihood (log10) = %.3f\n", best / log(10.0));
    return best;
}

// DescentGraph::get_likelihood (descent_graph.cc:150-156) of a host graph, on the device
double GPUMarkovChain::get_likelihood(DescentGraph& dg) {
    double v = 0.0;
    die_on(slk_dg_upload(chain, dg.get_internal_ptr()), "GPUMarkovChain::get_likelihood (upload)");
    die_on(slk_dg_likelihood(chain, &v), "GPUMarkovChain::get_likelihood");
    return v;
}

void GPUMarkovChain::upload(DescentGraph& dg) {
    die_on(slk_dg_upload(chain, dg.get_internal_ptr()), "GPUMarkovChain::upload");
}

void GPUMarkovChain::download(DescentGraph& dg) {
    die_on(slk_dg_download(chain, dg.get_internal_ptr()), "GPUMarkovChain::download");
}

double GPUMarkovChain::get_likelihood() {
    double v = 0.0;
    die_on(slk_dg_likelihood(chain, &v), "GPUMarkovChain::get_likelihood");
    return v;
}

// MarkovChain::step (markov_chain.cc:107-207): no progress bar, no trace; "only score the coldest chain"
void GPUMarkovChain::step(int start_iteration, int step_size) {
    if(!scoring_started) {
        trait_prob = calc_trait_prob();
        die_on(slk_lodscore_init(chain), "GPUMarkovChain::step (init)");
        scoring_started = true;
    }
    for(int i = start_iteration; i < start_iteration + step_size; ++i) {
        if(slk_sweep_is_lsampler(chain, (uint64_t) i, options.lsampler_prob))
            die_on(slk_lsampler_sweep(chain, (uint64_t) i), "GPUMarkovChain::step (L-sampler)");
        else
            die_on(slk_msampler_sweep(chain, (uint64_t) i), "GPUMarkovChain::step (M-sampler)");
        if(i < options.burnin) continue;
        if(temperature != 1.0) continue;
        if((i % options.scoring_period) == 0) die_on(slk_lodscore_accumulate(chain), "GPUMarkovChain::step (scoring)");
    }
}

LODscores* GPUMarkovChain::get_result() {
    if(temperature != 1.0) {
        fprintf(stderr, "error: only the coldest chain can be used!\n");
        abort();
    }
    LODscores* lod = new LODscores(map);
    std::vector<double> raw(lod->num_lodscores());
    int32_t count = 0;
    die_on(slk_lodscore_read(chain, raw.data(), &count), "GPUMarkovChain::get_result");
    lod->set_trait_prob(trait_prob);
    lod->set_count((unsigned int) count);
    if(count > 0) for(unsigned int i = 0; i < raw.size(); ++i) lod->set(i, raw[i]);
    return lod;
}

// markov_chain.cc:314-404 with every sampler and the scoring on the device: each iteration is an
// L-sweep with probability lsampler_prob, otherwise an M-sweep over the shuffled meioses
// (markov_chain.cc:332-349); scoring on iterations i >= burnin with i % scoring_period == 0
// (:351-361); optional CODA trace of the graph likelihood at the scored iterations (:362-370).
// The graph never leaves the device between sweeps (the reference's GPU chain copies it back and
// forth for every M-sweep, gpu_markov_chain.cc:1199-1210).
//
// The loop is split into begin / iterate / finish so that a caller can keep several replicate chains
// in flight on one device: every call only enqueues work on the chain's own stream (the CODA trace is
// the one exception), and a chain's draws are keyed by (seed, chain id, iteration), so interleaving
// chains changes nothing in their results.
void GPUMarkovChain::begin(DescentGraph& dg) {
    trait_prob = calc_trait_prob();
    printf("P(T) = %.5f\n", trait_prob / log(10.0));
    die_on(slk_dg_upload(chain, dg.get_internal_ptr()), "GPUMarkovChain::run (upload)");
    const bool need_fag = options.lsampler_prob < 1.0 || options.coda_logging;
    if(need_fag) {
        double lik = 0.0;
        die_on(slk_dg_likelihood(chain, &lik), "GPUMarkovChain::run (likelihood)");
        if(lik == LOG_ILLEGAL) {
            fprintf(stderr, "error: descent graph illegal pre-markov chain...\n");
            abort();
        }
    }
    coda = 0;
    if(options.coda_logging) {
        char buf[16];
        snprintf(buf, sizeof(buf), "%d", seq_num);
        std::string fname = options.coda_prefix + ".ped" + ped->get_id() + ".run" + std::string(buf);   // markov_chain.cc:84-91
        coda = fopen(fname.c_str(), "w");
        if(!coda) { fprintf(stderr, "error: could not open trace file '%s'\n", fname.c_str()); abort(); }
        fprintf(coda, "iteration likelihood\n");
        printf("opened trace file (%s)\n", fname.c_str());
    }
    die_on(slk_lodscore_init(chain), "GPUMarkovChain::run (init)");
    scoring_started = true;
}

void GPUMarkovChain::iterate(int i) {
    if(slk_sweep_is_lsampler(chain, (uint64_t) i, options.lsampler_prob))
        die_on(slk_lsampler_sweep(chain, (uint64_t) i), "GPUMarkovChain::run (L-sampler)");
    else
        die_on(slk_msampler_sweep(chain, (uint64_t) i), "GPUMarkovChain::run (M-sampler)");
    if(i < options.burnin) return;
    if((i % options.scoring_period) == 0) {
        if(coda) {
            double lik = 0.0;
            die_on(slk_dg_likelihood(chain, &lik), "GPUMarkovChain::run (trace)");
            if(lik == LOG_ILLEGAL) { fprintf(stderr, "error: descent graph illegal...\n"); abort(); }
            fprintf(coda, "%d\t%f\n", i + 1, lik);
        }
        die_on(slk_lodscore_accumulate(chain), "GPUMarkovChain::run (scoring)");
    }
}

LODscores* GPUMarkovChain::finish(DescentGraph& dg) {
    if(coda) { fclose(coda); coda = 0; }
    die_on(slk_dg_download(chain, dg.get_internal_ptr()), "GPUMarkovChain::run (download)");
    LODscores* lod = new LODscores(map);
    std::vector<double> raw(lod->num_lodscores());
    int32_t count = 0;
    die_on(slk_lodscore_read(chain, raw.data(), &count), "GPUMarkovChain::run (results)");
    lod->set_trait_prob(trait_prob);
    lod->set_count((unsigned int) count);
    for(unsigned int i = 0; i < raw.size(); ++i) lod->set(i, raw[i]);
    return lod;
}

LODscores* GPUMarkovChain::run(DescentGraph& dg) {
    begin(dg);
    const int total = options.iterations + options.burnin;
    for(int i = 0; i < total; ++i) iterate(i);
    return finish(dg);
}

// LinkageProgram::run_pedigree's replicate loop (linkage_program.cc:96-108) with up to `in_flight` replicates
// resident on the device at once.  One chain leaves most of a B200 idle during an M-sweep (its
// likelihood kernel keeps ~6 warps per SM busy, its chain kernel 8 SMs), so the replicates of a group
// are advanced one iteration at a time in turn, each on its own stream, and their kernels overlap.
// Results are merged in replicate order and are bit-identical to running the replicates one by one.
LODscores* run_replicates(Pedigree* ped, GeneticMap* map, PeelSequenceGenerator* psg, struct mcmc_options options, int in_flight) {
    if(in_flight < 1) in_flight = 1;
    LODscores* total = 0;
    for(int r0 = 0; r0 < options.mcmc_runs; r0 += in_flight) {
        const int n = std::min(in_flight, options.mcmc_runs - r0);
        std::vector<GPUMarkovChain*> chains;
        std::vector<DescentGraph*> dgs;
        for(int k = 0; k < n; ++k) {
            dgs.push_back(new DescentGraph(ped, map, options.sex_linked));
            chains.push_back(new GPUMarkovChain(ped, map, psg, options, r0 + k));
            chains[k]->sequential_imputation(*dgs[k], options.si_iterations);
        }
        for(int k = 0; k < n; ++k) chains[k]->begin(*dgs[k]);
        const int iters = options.iterations + options.burnin;
        for(int i = 0; i < iters; ++i)
            for(int k = 0; k < n; ++k) chains[k]->iterate(i);
        for(int k = 0; k < n; ++k) {
            LODscores* lod = chains[k]->finish(*dgs[k]);
            if(!total) total = lod;
            else { total->merge_results(lod); delete lod; }
            delete chains[k];
            delete dgs[k];
        }
    }
    return total;
}

}  // namespace swiftlink
