// job.cc -- one device's share of a `-R` job: the replicates of LinkageProgram::run_pedigree's loop
// (linkage_program.cc:96-108) that were placed on this device, each either a plain chain (MarkovChain::run,
// markov_chain.cc:314-404) or a Metropolis-coupled ladder (Mc3::run, mc3.cc:81-200), all of them resident on the
// device at once and advanced in turn -- every call only enqueues work on the chains' own streams, so their
// kernels overlap -- and the raw log-sum LOD accumulators of the finished replicates merged the way
// LODscores::merge_results does (lod_score.h:98-105).
//
// A job is what ONE process (one GPU) runs.  Across processes the replicates are dealt out round-robin
// (replicate r on rank r mod world) and the per-rank tables and swap counters are merged by small all-reduces
// (swiftlink_b200/run.py, swiftlink_b200/dist.py: NCCL over NVLink); a chain's draws are keyed by (seed, chain id,
// iteration), so the merged result does not depend on the number of GPUs.
#include "swiftlink_host.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace swiftlink {

ReplicateJob::ReplicateJob(Pedigree* ped, GeneticMap* map, PeelSequenceGenerator* psg, struct mcmc_options options,
                           const std::vector<int>& replicate_ids) :
    ped(ped), map(map), psg(psg), options(options), ids(replicate_ids), next_iteration(0) {
    ladder = options.mc3 && options.mc3_number_of_chains > 1;
    for(size_t k = 0; k < ids.size(); ++k) {
        if(ladder) {
            ladders.push_back(new Mc3(ped, map, psg, options, ids[k]));
            ladders[k]->start();
        }
        else {
            dgs.push_back(new DescentGraph(ped, map, options.sex_linked));
            chains.push_back(new GPUMarkovChain(ped, map, psg, options, ids[k]));
            chains[k]->sequential_imputation(*dgs[k], options.si_iterations);
        }
    }
    for(size_t k = 0; k < chains.size(); ++k) chains[k]->begin(*dgs[k]);
}

ReplicateJob::~ReplicateJob() {
    for(size_t k = 0; k < chains.size(); ++k) { delete chains[k]; delete dgs[k]; }
    for(size_t k = 0; k < ladders.size(); ++k) delete ladders[k];
}

int ReplicateJob::total_iterations() const { return options.burnin + options.iterations; }

// iterations next_iteration .. next_iteration + n - 1 of every replicate (a ladder advances in whole spurts of its
// exchange period: n is rounded down to a multiple of it); returns the iterations actually advanced
int ReplicateJob::advance(int n) {
    const int left = total_iterations() - next_iteration;
    if(n > left) n = left;
    if(n <= 0) return 0;
    if(ladder) {
        const int period = ladders.empty() ? 1 : ladders[0]->exchange_period();
        const int spurts = n / period;
        for(int s = 0; s < spurts; ++s) {
            // every ladder's chains first, then every ladder's swap test: the host waits (in exchange) while the
            // other ladders' spurts are still running
            for(size_t k = 0; k < ladders.size(); ++k) ladders[k]->enqueue_spurt();
            for(size_t k = 0; k < ladders.size(); ++k) ladders[k]->exchange();
        }
        next_iteration += spurts * period;
        return spurts * period;
    }
    for(int i = next_iteration; i < next_iteration + n; ++i)
        for(size_t k = 0; k < chains.size(); ++k) chains[k]->iterate(i);
    next_iteration += n;
    return n;
}

// merged table of the job's replicates so far (the caller deletes it); swap counters summed over the ladders
LODscores* ReplicateJob::results(std::vector<int>* swap_success, std::vector<int>* swap_failure) {
    LODscores* total = 0;
    const size_t n = ladder ? ladders.size() : chains.size();
    if(swap_success) swap_success->assign((size_t) std::max(options.mc3_number_of_chains, 1), 0);
    if(swap_failure) swap_failure->assign((size_t) std::max(options.mc3_number_of_chains, 1), 0);
    for(size_t k = 0; k < n; ++k) {
        LODscores* lod = ladder ? ladders[k]->cold_result() : chains[k]->get_result();
        if(!total) total = lod;
        else { total->merge_results(lod); delete lod; }
        if(ladder) {
            for(size_t i = 0; i < ladders[k]->get_swap_success().size(); ++i) {
                if(swap_success) (*swap_success)[i] += ladders[k]->get_swap_success()[i];
                if(swap_failure) (*swap_failure)[i] += ladders[k]->get_swap_failure()[i];
            }
        }
    }
    if(!total) total = new LODscores(map);
    return total;
}

}  // namespace swiftlink
