// mc3.cc -- Metropolis-coupled MCMC over device-resident chains.  Behavioural spec: the reference's
// Mc3 (mc3.h, mc3.cc:20-200), which is compiled but never reached (linkage_program.cc:169-170) and
// whose driver passes `i * spurts` as the first iteration of spurt i (mc3.cc:117; the burn-in test and
// the scoring period in MarkovChain::step then see the wrong iteration numbers).  Here iterations are
// numbered consecutively; everything else follows the reference:
//   ladder       T_0 = 1, T_i = 1 / (1 + 0.001 * 2^i), or the user's list        (mc3.cc:31-42)
//   heating      theta and minor allele frequency pulled towards 0.5            (genetic_map.cc:95-117)
//   start state  one sequential-imputation state per chain, on the cold map     (mc3.cc:91-102)
//   swap         pick adjacent chains r, r+1 uniformly; accept with probability
//                min(1, exp((L_r(g_{r+1}) + L_{r+1}(g_r)) - (L_r(g_r) + L_{r+1}(g_{r+1}))))  (mc3.cc:138-162)
//   result       the cold chain's LOD table                                      (mc3.cc:199)
// All chains of a ladder share one device, so the four likelihoods of a swap test are two
// slk_dg_likelihood calls before and two after an O(1) pointer exchange, which is undone on rejection.
#include "swiftlink_host.h"

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace swiftlink {

double mc3_temperature(int i, const struct mcmc_options& options) {
    if(!options.mc3) return 1.0;
    if(!options.mc3_temperatures.empty()) return options.mc3_temperatures[i];
    return i == 0 ? 1.0 : 1.0 / (1.0 + (0.001 * pow(2.0, i)));
}

Mc3::Mc3(Pedigree* ped, GeneticMap* map, PeelSequenceGenerator* psg, struct mcmc_options options, int sequence_num) :
    ped(ped), map(map), psg(psg), options(options), seq_num(sequence_num), period(10), spurts_done(0), rng(0) {
    if(options.mc3_number_of_chains < 1) { fprintf(stderr, "error: number of Markov chains must be greater than zero\n"); abort(); }
    if(!options.mc3_temperatures.empty() && (int) options.mc3_temperatures.size() != options.mc3_number_of_chains) {
        fprintf(stderr, "error: %d temperatures given for %d chains\n", (int) options.mc3_temperatures.size(), options.mc3_number_of_chains);
        abort();
    }
    for(int i = 0; i < options.mc3_number_of_chains; ++i) {
        const double t = mc3_temperature(i, options);
        fprintf(stderr, "Creating Markov chain %d, temperature = %.3f\n", i, t);
        // chain id: the replicate's ladder occupies ids seq * chains .. seq * chains + chains - 1
        chains.push_back(new GPUMarkovChain(ped, map, psg, options, sequence_num * options.mc3_number_of_chains + i, t));
    }
    swap_success.assign(chains.size(), 0);
    swap_failure.assign(chains.size(), 0);
}

Mc3::~Mc3() {
    for(size_t i = 0; i < chains.size(); ++i) delete chains[i];
}

// Mc3::run in four parts, so that a caller can keep several ladders in flight on one device (ReplicateJob, job.cc):
// enqueue_spurt only enqueues work on the chains' streams; exchange is where the host waits for the device.
void Mc3::start() {
    // start states (mc3.cc:91-102): sequential imputation with the cold chain's plan
    for(size_t j = 0; j < chains.size(); ++j) {
        DescentGraph tmp(ped, map, options.sex_linked);
        struct mcmc_options o = options;
        o.seed = options.seed + 7919ull * (j + 1);
        GPUMarkovChain starter(ped, map, psg, o, seq_num);
        starter.sequential_imputation(tmp, options.si_iterations);
        chains[j]->upload(tmp);
    }
    period = options.mc3_exchange_period;
    if(!options.mc3 || chains.size() == 1) period = 10;                   // mc3.cc:108-110
    spurts_done = 0;
    rng = HostRng(options.seed ^ (0x4d4333ull << 32) ^ (uint64_t) seq_num);
}

int Mc3::total_spurts() const { return (options.burnin + options.iterations) / period; }

void Mc3::enqueue_spurt() {
    for(size_t j = 0; j < chains.size(); ++j) chains[j]->step(spurts_done * period, period);
}

void Mc3::exchange() {
    ++spurts_done;
    if(!(options.mc3 && chains.size() > 1)) return;
    const int r = (int)((chains.size() - 1) * rng.uniform());
    slk_chain* a = chains[r]->get_chain();
    slk_chain* b = chains[r + 1]->get_chain();
    const double xx = chains[r]->get_likelihood(), yy = chains[r + 1]->get_likelihood();
    if(slk_dg_swap(a, b) != SLK_OK) { fprintf(stderr, "error: Mc3 swap: %s\n", slk_last_error()); abort(); }
    const double xy = chains[r]->get_likelihood(), yx = chains[r + 1]->get_likelihood();
    const double ratio = (xy + yx) - (xx + yy);
    const double u = rng.uniform();
    if(u == 0.0 || log(u) < std::min(0.0, ratio)) swap_success[r] += 1;
    else {
        swap_failure[r] += 1;
        if(slk_dg_swap(a, b) != SLK_OK) { fprintf(stderr, "error: Mc3 swap: %s\n", slk_last_error()); abort(); }
    }
}

LODscores* Mc3::finish() {
    if(options.mc3) {
        if(!options.exchange_filename.empty()) {
            FILE* ef = fopen(options.exchange_filename.c_str(), "w");
            if(ef) {
                for(int i = 0; i + 1 < (int) chains.size(); ++i)
                    fprintf(ef, "%d %g\n", i, swap_success[i] / (double)(swap_success[i] + swap_failure[i]));
                fclose(ef);
            }
        }
        for(int i = 0; i + 1 < (int) chains.size(); ++i)
            fprintf(stderr, "%d -- %d : %.3f (%d/%d)\n", i, i + 1,
                    swap_success[i] / (double)(swap_success[i] + swap_failure[i]), swap_success[i], swap_success[i] + swap_failure[i]);
    }
    if(!options.mc3 && chains.size() != 1) {                              // mc3.cc:191-197: independent chains, merged
        LODscores* tmp = chains[0]->get_result();
        for(size_t i = 1; i < chains.size(); ++i) { LODscores* o = chains[i]->get_result(); tmp->merge_results(o); delete o; }
        return tmp;
    }
    return chains[0]->get_result();
}

LODscores* Mc3::run() {
    start();
    const int spurts = total_spurts();
    for(int i = 0; i < spurts; ++i) {
        enqueue_spurt();
        exchange();
    }
    return finish();
}

}  // namespace swiftlink
