// elod.cc -- expected LOD by simulation, on the device.  Same set-up, messages and result as the reference's
// Elod (elod.h:33-107, elod.cc:19-85); the replicate loop (elod.cc:55-66: LocusSampler::start_from,
// two DescentGraph::copy_locus, Peeler::process) is one batched call, slk_elod_run.
#include "swiftlink_host.h"

#include <cmath>
#include <cstdlib>

namespace swiftlink {

Elod::Elod(const char* pedfile, struct mcmc_options opt) :
    dm(opt.elod_frequency, opt.elod_penetrance, opt.sex_linked), map1(1), map2(1), options(opt) {
    // "setup a fake map" (elod.h:40-66): marker, trait, marker; and the two markers alone
    Snp s0("marker", 0.0);
    s0.set_minor_freq(0.5);
    Snp s1("trait", options.elod_marker_separation / 2);
    s1.set_minor_freq(options.elod_frequency);
    Snp s2("marker", options.elod_marker_separation);
    s2.set_minor_freq(0.5);
    map1.add(s0); map1.add(s1); map1.add(s2);
    map1.add_theta(map1.haldane(options.elod_marker_separation / 2));
    map1.add_theta(map1.haldane(options.elod_marker_separation / 2));
    map2.add(s0); map2.add(s2);
    map2.add_theta(map2.haldane(options.elod_marker_separation));
    if(!map1.sanity_check() || !map2.sanity_check()) exit(1);

    if(!parse_pedigree_file(pedfile, pedigrees, dm, map1, true)) exit(1);

    // "kill any genotype information read in, add UNTYPED for both loci in the fake map, then copy across
    // the probabilities from the disease trait" (elod.h:73-90)
    for(size_t i = 0; i < pedigrees.size(); ++i) {
        for(unsigned j = 0; j < pedigrees[i].num_members(); ++j) {
            Person* p = pedigrees[i].get_by_index(j);
            p->clear_genotypes();
            p->add_genotype(UNTYPED); p->add_genotype(UNTYPED); p->add_genotype(UNTYPED);
            p->populate_trait_prob_cache(map1, options.sex_linked);
            p->copy_disease_probs(1);
            if(options.affected_only && !p->isaffected()) p->make_unknown_affection(dm);
        }
    }
}

static void die_on(int rc, const char* what) {
    if(rc != SLK_OK) { fprintf(stderr, "error: %s: %s\n", what, slk_last_error()); abort(); }
}

double Elod::run() {
    elods.clear();
    fprintf(stderr, "\nELOD parameters:\n\tpenetrance = %.2f:%.2f:%.2f\n\tseparation = %.2f\n\ttrait freq = %.2e\n"
                    "\treplicates = %d\n\tsex-linked = %s\n\n",
            options.elod_penetrance[0], options.elod_penetrance[1], options.elod_penetrance[2], options.elod_marker_separation,
            options.elod_frequency, options.elod_replicates, options.sex_linked ? "true" : "false");

    for(size_t i = 0; i < pedigrees.size(); ++i) {
        PeelSequenceGenerator psg(&pedigrees[i], &map1, options.sex_linked, options.verbose, options.seed);
        psg.build_peel_sequence(options.peelopt_iterations);

        // sampler plan on the three-locus map (genotype priors with the resolved founder flags, the trait
        // locus's prior = disease probabilities), trait plan on the two-marker map
        FlatProblem f1, f2;
        flatten_problem(pedigrees[i], map1, psg, options.sex_linked, f1, true, 1);
        // the two-locus problem shares everything but the map: drop the middle locus
        flatten_problem(pedigrees[i], map1, psg, options.sex_linked, f2, true, -1);
        const int N = (int) pedigrees[i].num_members();
        {
            std::vector<uint8_t> g(2 * (size_t) N), e(2 * (size_t) N);
            for(int p = 0; p < N; ++p) { g[2 * p] = f2.genotypes[3 * p]; g[2 * p + 1] = f2.genotypes[3 * p + 2]; }
            for(int p = 0; p < N; ++p) { e[p] = f2.elimination[p]; e[N + p] = f2.elimination[2 * N + p]; }
            f2.genotypes = g; f2.elimination = e;
            f2.marker_prob.resize(8); f2.marker_xprob.resize(8); f2.minor_freq.resize(2);
            for(int k = 0; k < 4; ++k) {
                f2.marker_prob[k] = map2.get_prob(0, (enum phased_trait) k, false); f2.marker_prob[4 + k] = map2.get_prob(1, (enum phased_trait) k, false);
                f2.marker_xprob[k] = map2.get_prob(0, (enum phased_trait) k, true); f2.marker_xprob[4 + k] = map2.get_prob(1, (enum phased_trait) k, true);
            }
            f2.minor_freq[0] = map2.get_minor(0); f2.minor_freq[1] = map2.get_minor(1);
            f2.theta.assign(1, map2.get_theta(0));
            f2.partial_theta.assign(1, map2.get_theta_partial_raw(0));
            slk_problem& d = f2.desc;
            d.n_markers = 2; d.n_lod = (int) map2.get_lodscore_count();
            d.genotypes = f2.genotypes.data(); d.elimination = f2.elimination.data();
            d.marker_prob = f2.marker_prob.data(); d.marker_xprob = f2.marker_xprob.data(); d.minor_freq = f2.minor_freq.data();
            d.theta = f2.theta.data(); d.partial_theta = f2.partial_theta.data();
        }
        slk_plan *p1 = 0, *p2 = 0;
        die_on(slk_plan_create(&f1.desc, options.device, &p1), "Elod: sampler plan");
        die_on(slk_plan_create(&f2.desc, options.device, &p2), "Elod: trait plan");
        double trait_prob = 0.0, log_sum = 0.0;
        int64_t count = 0;
        die_on(slk_trait_likelihood(p2, &trait_prob), "Elod: trait likelihood");
        die_on(slk_elod_run(p1, p2, options.seed, (uint32_t) i, options.elod_replicates, &log_sum, &count, 0), "Elod: replicates");
        slk_plan_destroy(p1);
        slk_plan_destroy(p2);
        // LODscores::get(0, 0) (lod_score.h:86-88)
        elods.push_back((log_sum - log((double) count) - trait_prob) / log(10.0));
    }

    double total_elod = 0.0;
    fprintf(stderr, "\n%10s |%12s\n", "Pedigree", "ELOD");
    fprintf(stderr, "-----------|------------\n");
    for(size_t i = 0; i < elods.size(); ++i) {
        fprintf(stderr, "%10d |%12f\n", (int) i, elods[i]);
        total_elod += elods[i];
    }
    fprintf(stderr, "%10s |%12f\n", "Total", total_elod);
    return total_elod;
}

}  // namespace swiftlink
