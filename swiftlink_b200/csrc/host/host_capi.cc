// host_capi.cc -- C shims declared in include/swiftlink_b200_host.h.
#include "swiftlink_b200_host.h"
#include "swiftlink_host.h"

#include <cmath>
#include <cstring>

using namespace swiftlink;

struct slk_host {
    GeneticMap map;
    DiseaseModel dm;
    std::vector<Pedigree> peds;
    Pedigree* ped;
    PeelSequenceGenerator* psg;
    FlatProblem flat;
    bool flat_valid;
    explicit slk_host(int nlod) : map(nlod), ped(0), psg(0), flat_valid(false) {}
};

extern "C" {

slk_host* slk_host_open(const char* pedfile, const char* mapfile, const char* datfile, int force_sex_linked, int n_lod) {
    slk_host* h = new slk_host(n_lod);
    if(!read_and_check_input(pedfile, mapfile, datfile, force_sex_linked != 0, h->map, h->dm, h->peds)) { delete h; return 0; }
    if(h->peds.size() != 1) {
        fprintf(stderr, "error: expected exactly one pedigree, read %d\n", (int) h->peds.size());
        delete h; return 0;
    }
    h->ped = &h->peds[0];
    return h;
}

void slk_host_close(slk_host* h) {
    if(!h) return;
    delete h->psg;
    delete h;
}

void slk_host_dims(const slk_host* h, int32_t* out) {
    out[0] = (int32_t) h->ped->num_members();
    out[1] = (int32_t) h->ped->num_founders();
    out[2] = (int32_t) h->map.num_markers();
    out[3] = (int32_t) h->map.get_lodscore_count();
    out[4] = h->dm.is_sexlinked() ? 1 : 0;
}

void slk_host_person_table(const slk_host* h, int32_t* mother, int32_t* father, int32_t* sex, int32_t* affection,
                           int32_t* typed, double* disease_prob) {
    for(unsigned i = 0; i < h->ped->num_members(); ++i) {
        const Person* p = h->ped->get_by_index(i);
        mother[i] = p->isfounder() ? -1 : (int32_t) p->get_maternalid();
        father[i] = p->isfounder() ? -1 : (int32_t) p->get_paternalid();
        sex[i] = (int32_t) p->get_sex();
        affection[i] = (int32_t) p->get_affection();
        typed[i] = p->istyped() ? 1 : 0;
        for(int j = 0; j < 4; ++j) disease_prob[i * 4 + j] = p->get_disease_prob((enum phased_trait) j);
    }
}

int slk_host_person_name(const slk_host* h, int i, char* buf, int cap) {
    std::string s = h->ped->get_by_index(i)->get_id();
    snprintf(buf, cap, "%s", s.c_str());
    return (int) s.size();
}

int slk_host_marker_name(const slk_host* h, int l, char* buf, int cap) {
    std::string s = h->map.get_name(l);
    snprintf(buf, cap, "%s", s.c_str());
    return (int) s.size();
}

void slk_host_genotypes(const slk_host* h, int32_t* out) {
    const unsigned M = h->map.num_markers();
    for(unsigned i = 0; i < h->ped->num_members(); ++i)
        for(unsigned l = 0; l < M; ++l) out[i * M + l] = (int32_t) h->ped->get_by_index(i)->get_marker(l);
}

void slk_host_marker_trait_prob(const slk_host* h, double* out) {
    const unsigned M = h->map.num_markers();
    for(unsigned i = 0; i < h->ped->num_members(); ++i)
        for(unsigned l = 0; l < M; ++l)
            for(int j = 0; j < 4; ++j)
                out[((size_t) i * M + l) * 4 + j] = h->ped->get_by_index(i)->get_trait_probability(l, (enum phased_trait) j);
}

void slk_host_map_table(const slk_host* h, double* gdist, double* minor, double* prob, double* xprob,
                        double* theta, double* partial) {
    const unsigned M = h->map.num_markers();
    for(unsigned l = 0; l < M; ++l) {
        gdist[l] = h->map[l].get_g_distance();
        minor[l] = h->map[l].minor();
        for(int j = 0; j < 4; ++j) {
            prob[l * 4 + j] = h->map.get_prob(l, (enum phased_trait) j, false);
            xprob[l * 4 + j] = h->map.get_prob(l, (enum phased_trait) j, true);
        }
    }
    for(unsigned l = 0; l + 1 < M; ++l) { theta[l] = h->map.get_theta(l); partial[l] = h->map.get_theta_partial_raw(l); }
}

void slk_host_disease_model(const slk_host* h, double* out) {
    out[0] = h->dm.get_freq();
    for(int i = 0; i < 3; ++i) out[1 + i] = h->dm.get_penetrance((enum unphased_trait) i);
}

static void need_psg(slk_host* h, uint64_t seed) {
    if(!h->psg) h->psg = new PeelSequenceGenerator(h->ped, &h->map, h->dm.is_sexlinked(), false, seed);
}

void slk_host_elim_masks(slk_host* h, int32_t* out) {
    need_psg(h, 20261017);
    const unsigned M = h->map.num_markers(), N = h->ped->num_members();
    for(unsigned l = 0; l < M; ++l) for(unsigned i = 0; i < N; ++i) out[l * N + i] = h->psg->get_elimination().mask(l, i);
}

void slk_host_build_peel(slk_host* h, int iterations, uint64_t seed) {
    delete h->psg;
    h->psg = 0;
    need_psg(h, seed);
    h->psg->build_peel_sequence((unsigned) iterations);
    h->flat_valid = false;
}

int slk_host_set_peel(slk_host* h, const uint32_t* seq) {
    delete h->psg;
    h->psg = 0;
    need_psg(h, 20261017);
    std::vector<unsigned int> s(seq, seq + h->ped->num_members());
    h->flat_valid = false;
    return h->psg->set_peel_sequence(s) ? 1 : 0;
}

int slk_host_num_ops(const slk_host* h) { return h->psg ? (int) h->psg->get_peel_order().size() : 0; }
uint32_t slk_host_peel_cost(const slk_host* h) { return h->psg ? h->psg->get_peeling_cost() : 0; }

void slk_host_op_info(const slk_host* h, int i, int32_t* info, int32_t* cutset, int32_t* previous, int32_t* children) {
    const PeelOperation& op = h->psg->get_peel_order()[i];
    info[0] = (int32_t) op.get_type();
    info[1] = (int32_t) op.get_peelnode();
    info[2] = (int32_t) op.get_cutset_size();
    info[3] = (int32_t) op.get_prevfunctions().size();
    info[4] = (int32_t) op.get_children().size();
    for(unsigned j = 0; j < op.get_cutset_size(); ++j) cutset[j] = (int32_t) op.get_cutnode(j);
    for(size_t j = 0; j < op.get_prevfunctions().size(); ++j) previous[j] = (int32_t) op.get_prevfunctions()[j];
    for(size_t j = 0; j < op.get_children().size(); ++j) children[j] = (int32_t) op.get_children()[j];
}

int slk_host_random_descentgraph(slk_host* h, uint64_t seed, int32_t* dg) {
    DescentGraph d(h->ped, &h->map, h->dm.is_sexlinked());
    HostRng rng(seed);
    if(!d.random_descentgraph(rng)) return 0;
    memcpy(dg, d.get_internal_ptr(), d.get_internal_size());
    return 1;
}

const slk_problem* slk_host_problem(slk_host* h) {
    if(!h->psg) return 0;
    if(!h->flat_valid) {
        flatten_problem(*h->ped, h->map, *h->psg, h->dm.is_sexlinked(), h->flat);
        h->flat_valid = true;
    }
    return &h->flat.desc;
}

int slk_host_write_results(slk_host* h, const char* filename, const double* lod) {
    LODscores ls(&h->map);
    // feed already-normalised values through the writer: count = 1, trait_prob = 0, ln10 scaling undone
    ls.set_count(1);
    ls.set_trait_prob(0.0);
    for(unsigned i = 0; i < ls.num_lodscores(); ++i) ls.set(i, lod[i] * log(10.0));
    std::vector<LODscores*> all(1, &ls);
    return write_linkage_results(&h->map, filename, all, false) ? 1 : 0;
}

int slk_host_run_chain(slk_host* h, int device, uint64_t seed, uint32_t chain_id, int burnin, int iterations,
                       int scoring_period, double lsampler_prob, int32_t* dg, double* lod_out, double* trait_prob_out) {
    if(!h->psg) return SLK_ERR_INVALID;
    struct mcmc_options o;
    o.burnin = burnin; o.iterations = iterations; o.scoring_period = scoring_period;
    o.lsampler_prob = lsampler_prob; o.sex_linked = h->dm.is_sexlinked(); o.seed = seed; o.device = device;
    DescentGraph d(h->ped, &h->map, h->dm.is_sexlinked());
    memcpy(d.get_internal_ptr(), dg, d.get_internal_size());
    GPUMarkovChain chain(h->ped, &h->map, h->psg, o, (int) chain_id);
    LODscores* lod = chain.run(d);
    memcpy(dg, d.get_internal_ptr(), d.get_internal_size());
    const unsigned n = h->map.get_lodscore_count();
    for(unsigned l = 0; l + 1 < h->map.num_markers(); ++l) for(unsigned k = 0; k < n; ++k) lod_out[l * n + k] = lod->get(l, k);
    if(trait_prob_out) *trait_prob_out = lod->get_trait_prob();
    delete lod;
    return SLK_OK;
}

int slk_host_run_replicates(slk_host* h, int device, uint64_t seed, int runs, int in_flight, int burnin, int iterations,
                            int scoring_period, double lsampler_prob, int si_iterations, double* lod_out) {
    if(!h->psg || runs < 1) return SLK_ERR_INVALID;
    struct mcmc_options o;
    o.burnin = burnin; o.iterations = iterations; o.scoring_period = scoring_period; o.si_iterations = si_iterations;
    o.lsampler_prob = lsampler_prob; o.sex_linked = h->dm.is_sexlinked(); o.seed = seed; o.device = device; o.mcmc_runs = runs;
    LODscores* lod = run_replicates(h->ped, &h->map, h->psg, o, in_flight);
    const unsigned n = h->map.get_lodscore_count();
    for(unsigned l = 0; l + 1 < h->map.num_markers(); ++l) for(unsigned k = 0; k < n; ++k) lod_out[l * n + k] = lod->get(l, k);
    delete lod;
    return SLK_OK;
}

int slk_host_run_mc3(slk_host* h, int device, uint64_t seed, uint32_t chain_id, int n_chains, int exchange_period,
                     const double* temperatures, int burnin, int iterations, int scoring_period, double lsampler_prob,
                     int si_iterations, double* lod_out, int32_t* swap_success, int32_t* swap_failure) {
    if(!h->psg) return SLK_ERR_INVALID;
    struct mcmc_options o;
    o.burnin = burnin; o.iterations = iterations; o.scoring_period = scoring_period; o.si_iterations = si_iterations;
    o.lsampler_prob = lsampler_prob; o.sex_linked = h->dm.is_sexlinked(); o.seed = seed; o.device = device;
    o.mc3 = true; o.mc3_number_of_chains = n_chains; o.mc3_exchange_period = exchange_period;
    if(temperatures) o.mc3_temperatures.assign(temperatures, temperatures + n_chains);
    Mc3 mc3(h->ped, &h->map, h->psg, o, (int) chain_id);
    LODscores* lod = mc3.run();
    const unsigned n = h->map.get_lodscore_count();
    for(unsigned l = 0; l + 1 < h->map.num_markers(); ++l) for(unsigned k = 0; k < n; ++k) lod_out[l * n + k] = lod->get(l, k);
    for(int i = 0; i < n_chains; ++i) {
        if(swap_success) swap_success[i] = mc3.get_swap_success()[i];
        if(swap_failure) swap_failure[i] = mc3.get_swap_failure()[i];
    }
    delete lod;
    return SLK_OK;
}

double slk_host_elod(const char* pedfile, double frequency, const double* penetrance, double separation, int replicates,
                     int sex_linked, int affected_only, int peel_iterations, uint64_t seed, int device,
                     double* per_pedigree, int cap) {
    struct mcmc_options o;
    o.elod = true; o.elod_frequency = frequency; o.elod_marker_separation = separation; o.elod_replicates = replicates;
    for(int i = 0; i < 3; ++i) o.elod_penetrance[i] = penetrance[i];
    o.sex_linked = sex_linked != 0; o.affected_only = affected_only != 0; o.peelopt_iterations = peel_iterations;
    o.seed = seed; o.device = device;
    Elod e(pedfile, o);
    const double total = e.run();
    for(int i = 0; i < (int) e.per_pedigree().size() && i < cap; ++i) per_pedigree[i] = e.per_pedigree()[i];
    return total;
}

struct slk_host_job {
    slk_host* host;
    ReplicateJob* job;
};

slk_host_job* slk_host_job_create(slk_host* h, int device, uint64_t seed, const int32_t* replicate_ids, int n_replicates,
                                  int mc3_chains, int exchange_period, const double* temperatures, int burnin, int iterations,
                                  int scoring_period, double lsampler_prob, int si_iterations) {
    if(!h || !h->psg || n_replicates < 0 || (n_replicates > 0 && !replicate_ids)) return 0;
    struct mcmc_options o;
    o.burnin = burnin; o.iterations = iterations; o.scoring_period = scoring_period; o.si_iterations = si_iterations;
    o.lsampler_prob = lsampler_prob; o.sex_linked = h->dm.is_sexlinked(); o.seed = seed; o.device = device;
    o.mc3 = mc3_chains > 1; o.mc3_number_of_chains = mc3_chains > 1 ? mc3_chains : 1; o.mc3_exchange_period = exchange_period;
    if(temperatures && mc3_chains > 1) o.mc3_temperatures.assign(temperatures, temperatures + mc3_chains);
    slk_host_job* j = new slk_host_job();
    j->host = h;
    j->job = new ReplicateJob(h->ped, &h->map, h->psg, o, std::vector<int>(replicate_ids, replicate_ids + n_replicates));
    return j;
}

int slk_host_job_advance(slk_host_job* j, int n_iterations) {
    return j ? j->job->advance(n_iterations) : 0;
}

int slk_host_job_results(slk_host_job* j, double* raw, int32_t* count, double* trait_prob, int32_t* swap_success, int32_t* swap_failure,
                         int n_swap) {
    if(!j) return SLK_ERR_INVALID;
    std::vector<int> ok, bad;
    LODscores* lod = j->job->results(&ok, &bad);
    if(raw) for(unsigned i = 0; i < lod->num_lodscores(); ++i) raw[i] = lod->get_count() > 0 ? lod->get_raw(i) : LOG_ZERO;
    if(count) *count = (int32_t) lod->get_count();
    if(trait_prob) *trait_prob = lod->get_trait_prob();
    for(int i = 0; i < n_swap; ++i) {
        if(swap_success) swap_success[i] = i < (int) ok.size() ? ok[i] : 0;
        if(swap_failure) swap_failure[i] = i < (int) bad.size() ? bad[i] : 0;
    }
    delete lod;
    return SLK_OK;
}

void slk_host_job_destroy(slk_host_job* j) {
    if(!j) return;
    delete j->job;
    delete j;
}

double slk_host_mc3_temperature(int chain_index, int n_chains, const double* temperatures) {
    struct mcmc_options o;
    o.mc3 = true; o.mc3_number_of_chains = n_chains;
    if(temperatures) o.mc3_temperatures.assign(temperatures, temperatures + n_chains);
    return mc3_temperature(chain_index, o);
}

}  // extern "C"
