/*
 * ref_driver.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * A thin extern "C" dump/bench driver around the UNMODIFIED reference sources, which the
 * recipe in oracle/Makefile compiles where they lie under /root/reference/src (nothing is
 * copied into this repository).  The result, oracle/_ref/libswiftref.so, is used by
 *   - tests/ (as the checker that pins oracle/peel_oracle.c and the CUDA path),
 *   - tests/golden/make_golden.py (to generate the committed fixtures),
 *   - bench.py --impl reference and the cpu_baseline leg (kind = "reference").
 *
 * The reference keeps its state in private members; this file includes the reference
 * headers with `private`/`protected` re-defined so the driver can read peel matrices,
 * index tables and elimination masks directly.  Access specifiers do not change GCC's
 * class layout, so the objects compiled from the pristine sources stay ABI-compatible.
 *
 * Reference entry points exercised here (file:line under /root/reference/src):
 *   Program::read_and_check_input      program.cc:18-63
 *   PeelSequenceGenerator              peel_sequence_generator.cc:225 (finalise_peel_order)
 *   LocusSampler::step / set_locus     locus_sampler2.cc:128-181
 *   SamplerRfunction::evaluate/sample  rfunction.cc:172, sampler_rfunction.cc:159
 *   Peeler::process / calc_trait_prob  peeler.cc:65-103
 *   MarkovChain::run                   markov_chain.cc:314
 *   SequentialImputation::parallel_run sequential_imputation.cc:67
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <map>
#include <queue>
#include <limits>
#include <sstream>
#include <fstream>
#include <iostream>
#include <iomanip>
#include <iterator>
#include <algorithm>
#include <omp.h>
#include <istream>
#include <math.h>
#include <float.h>
#include <stdio.h>
#include <unistd.h>
#include <fcntl.h>

#include <gsl/gsl_rng.h>

#define private public
#define protected public
#define class struct        /* members before the first access specifier are private in a class */
#include "types.h"
#include "genetic_map.h"
#include "disease_model.h"
#include "person.h"
#include "pedigree.h"
#include "map_parser.h"
#include "linkage_parser.h"
#include "pedigree_parser.h"
#include "descent_graph.h"
#include "elimination.h"
#include "peeling.h"
#include "peel_matrix.h"
#include "peel_sequence_generator.h"
#include "rfunction.h"
#include "sampler_rfunction.h"
#include "trait_rfunction.h"
#include "locus_sampler2.h"
#include "peeler.h"
#include "lod_score.h"
#include "markov_chain.h"
#include "sequential_imputation.h"
#include "meiosis_sampler.h"
#include "elod.h"
#include "random.h"
#include "omp_facade.h"
#ifdef REF_WITH_GPU
#include "gpu_lodscores.h"
#endif
#undef private
#undef protected
#undef class

extern gsl_rng** r;          /* random.cc:14-15, one stream per OpenMP thread */

namespace {

struct RefCtx {
    struct mcmc_options opt;
    GeneticMap map;
    DiseaseModel dm;
    std::vector<Pedigree> peds;
    Pedigree* ped;
    PeelSequenceGenerator* psg;
    DescentGraph* dg;
    LocusSampler* ls;
    LODscores* lod;
    Peeler* peeler;
    MeiosisSampler* ms;

    RefCtx(int nlod) : opt(), map(nlod), dm(), peds(), ped(0), psg(0), dg(0), ls(0), lod(0), peeler(0), ms(0) {}
};

bool rng_ready = false;
int rng_threads = 0;

/* silence the reference's chatter on stdout/stderr while a call runs */
struct Quiet {
    int so, se, active;
    Quiet(bool on) : so(-1), se(-1), active(on) {
        if(!active) return;
        fflush(stdout); fflush(stderr);
        so = dup(1); se = dup(2);
        int dn = open("/dev/null", O_WRONLY);
        dup2(dn, 1); dup2(dn, 2); close(dn);
    }
    ~Quiet() {
        if(!active) return;
        fflush(stdout); fflush(stderr);
        dup2(so, 1); dup2(se, 2); close(so); close(se);
    }
};

bool quiet_mode() { return getenv("SLK_REF_VERBOSE") == 0; }

void drop_samplers(RefCtx* c) {
    delete c->ls;     c->ls = 0;
    delete c->peeler; c->peeler = 0;
    delete c->lod;    c->lod = 0;
    delete c->ms;     c->ms = 0;
}

void need_samplers(RefCtx* c) {
    if(!c->psg) { fprintf(stderr, "ref_driver: peel sequence not built\n"); abort(); }
    if(!c->ls)     c->ls = new LocusSampler(c->ped, &c->map, c->psg, 0, c->dm.is_sexlinked());
    if(!c->lod)    c->lod = new LODscores(&c->map);
    if(!c->peeler) c->peeler = new Peeler(c->ped, &c->map, c->psg, c->lod, c->dm.is_sexlinked());
}

} // namespace

extern "C" {

/* ---- lifecycle ------------------------------------------------------------------- */

void ref_set_threads(int n) {
    if(rng_ready && n > rng_threads) {
        fprintf(stderr, "ref_driver: call ref_set_threads before ref_seed (rng has %d streams)\n", rng_threads);
        abort();
    }
    omp_set_num_threads(n);
}

void ref_seed(unsigned int seed) {
    if(!rng_ready) {
        init_random();
        rng_ready = true;
        rng_threads = get_max_threads();
    }
    for(int i = 0; i < rng_threads; ++i) {
        gsl_rng_set(r[i], seed + 7919u * (unsigned) i);
    }
    srand(seed);   /* libc rand(): random_shuffle and elimination.cc:275,286 */
}

/* program.cc:18-63, same order: map -> dat -> (force X) -> map sanity -> ped */
void* ref_open(const char* pedfile, const char* mapfile, const char* datfile, int force_sexlinked, int lodscores) {
    Quiet q(quiet_mode());
    RefCtx* c = new RefCtx(lodscores);
    c->opt.lodscores = lodscores;

    MapParser mp(mapfile, c->map);
    if(!mp.parse()) { delete c; return 0; }

    LinkageParser lp(datfile, c->map, c->dm);
    if(!lp.parse()) { delete c; return 0; }

    if(force_sexlinked) c->dm.set_sexlinked(true);

    if(!c->map.sanity_check()) { delete c; return 0; }

    PedigreeParser pp(pedfile, c->peds, c->dm, c->map);
    if(!pp.parse()) { delete c; return 0; }

    if(c->peds.size() != 1) {
        fprintf(stderr, "ref_driver: expected exactly one pedigree, read %d\n", (int) c->peds.size());
        delete c; return 0;
    }
    c->ped = &c->peds[0];
    if(c->map.num_markers() != c->ped->num_markers()) { delete c; return 0; }

    c->opt.sex_linked = c->dm.is_sexlinked();
    c->dg = new DescentGraph(c->ped, &c->map, c->dm.is_sexlinked());
    return c;
}

void ref_close(void* h) {
    RefCtx* c = (RefCtx*) h;
    drop_samplers(c);
    delete c->psg;
    delete c->dg;
    delete c;
}

/* out[0..5] = N, F, M, lodscores-per-interval, sex_linked, leaves */
void ref_dims(void* h, int* out) {
    RefCtx* c = (RefCtx*) h;
    out[0] = c->ped->num_members();
    out[1] = c->ped->num_founders();
    out[2] = c->map.num_markers();
    out[3] = c->map.get_lodscore_count();
    out[4] = c->dm.is_sexlinked() ? 1 : 0;
    out[5] = c->ped->num_leaves();
}

/* ---- input tables ---------------------------------------------------------------- */

/* per person: mother, father (-1 unknown), sex (types.h:25-29), affection (:31-35), typed flag,
   disease_prob[4] in phased_trait order UU,AA,AU,UA (person.cc:85-119) */
void ref_person_table(void* h, int* mother, int* father, int* sex, int* affection, int* typed, double* disease_prob) {
    RefCtx* c = (RefCtx*) h;
    for(unsigned i = 0; i < c->ped->num_members(); ++i) {
        Person* p = c->ped->get_by_index(i);
        mother[i] = p->isfounder() ? -1 : (int) p->get_maternalid();
        father[i] = p->isfounder() ? -1 : (int) p->get_paternalid();
        sex[i] = (int) p->get_sex();
        affection[i] = (int) p->get_affection();
        typed[i] = p->istyped() ? 1 : 0;
        for(int j = 0; j < 4; ++j) disease_prob[i*4 + j] = p->get_disease_prob(static_cast<enum phased_trait>(j));
    }
}

int ref_person_name(void* h, int i, char* buf, int cap) {
    RefCtx* c = (RefCtx*) h;
    std::string s = c->ped->get_by_index(i)->get_id();
    snprintf(buf, cap, "%s", s.c_str());
    return (int) s.size();
}

/* genotypes[N][M], codes from genotype.h:26-31 (UNTYPED=0, HETERO=1, HOMOZ_A=2, HOMOZ_B=3) */
void ref_genotypes(void* h, int* out) {
    RefCtx* c = (RefCtx*) h;
    unsigned M = c->map.num_markers();
    for(unsigned i = 0; i < c->ped->num_members(); ++i) {
        Person* p = c->ped->get_by_index(i);
        for(unsigned l = 0; l < M; ++l) out[i*M + l] = (int) p->get_marker(l);
    }
}

/* marker trait probabilities [N][M][4] (person.cc:224-245) */
void ref_marker_trait_prob(void* h, double* out) {
    RefCtx* c = (RefCtx*) h;
    unsigned M = c->map.num_markers();
    for(unsigned i = 0; i < c->ped->num_members(); ++i) {
        Person* p = c->ped->get_by_index(i);
        for(unsigned l = 0; l < M; ++l)
            for(int j = 0; j < 4; ++j)
                out[(i*M + l)*4 + j] = p->get_trait_probability(l, static_cast<enum phased_trait>(j));
    }
}

/* gdist[M] (Morgans), minor[M], prob[M][4], xprob[M][4], theta[M-1], partial[M-1] */
void ref_map_table(void* h, double* gdist, double* minor, double* prob, double* xprob, double* theta, double* partial) {
    RefCtx* c = (RefCtx*) h;
    unsigned M = c->map.num_markers();
    for(unsigned l = 0; l < M; ++l) {
        gdist[l] = c->map[l].get_g_distance();
        minor[l] = c->map[l].minor();
        for(int j = 0; j < 4; ++j) {
            prob[l*4 + j]  = c->map.get_prob(l, static_cast<enum phased_trait>(j), false);
            xprob[l*4 + j] = c->map.get_prob(l, static_cast<enum phased_trait>(j), true);
        }
    }
    for(unsigned l = 0; l + 1 < M; ++l) {
        theta[l] = c->map.get_theta(l);
        partial[l] = c->map.get_theta_partial_raw(l);
    }
}

int ref_marker_name(void* h, int l, char* buf, int cap) {
    RefCtx* c = (RefCtx*) h;
    std::string s = c->map.get_name(l);
    snprintf(buf, cap, "%s", s.c_str());
    return (int) s.size();
}

/* disease model: out[0]=freq, out[1..3]=penetrance (disease_model.h:15-16) */
void ref_disease_model(void* h, double* out) {
    RefCtx* c = (RefCtx*) h;
    out[0] = c->dm.get_freq();
    for(int i = 0; i < 3; ++i) out[1+i] = c->dm.get_penetrance(static_cast<enum unphased_trait>(i));
}

/* ---- peel plan ------------------------------------------------------------------- */

/* fresh generator; search as in linkage_program.cc:136-137 */
void ref_build_peel(void* h, int iterations) {
    RefCtx* c = (RefCtx*) h;
    Quiet q(quiet_mode());
    drop_samplers(c);
    delete c->psg;
    c->psg = new PeelSequenceGenerator(c->ped, &c->map, c->dm.is_sexlinked(), false);
    c->psg->build_peel_sequence(iterations);
}

/* fresh generator; force a given elimination order (peel_sequence_generator.cc:225) */
int ref_set_peel(void* h, const unsigned int* seq) {
    RefCtx* c = (RefCtx*) h;
    Quiet q(quiet_mode());
    drop_samplers(c);
    delete c->psg;
    c->psg = new PeelSequenceGenerator(c->ped, &c->map, c->dm.is_sexlinked(), false);
    std::vector<unsigned int> s(seq, seq + c->ped->num_members());
    if(!c->psg->is_legit(s)) return 0;
    c->psg->finalise_peel_order(s);
    return 1;
}

/* elimination masks [M][N], bits AA=8 AB=4 BA=2 BB=1 (genotype.h:18-24) */
void ref_elim_masks(void* h, int* out) {
    RefCtx* c = (RefCtx*) h;
    if(!c->psg) { fprintf(stderr, "ref_driver: peel sequence not built\n"); abort(); }
    unsigned M = c->map.num_markers(), N = c->ped->num_members();
    for(unsigned l = 0; l < M; ++l)
        for(unsigned i = 0; i < N; ++i)
            out[l*N + i] = c->psg->ge.possible_genotypes[l][i];
}

int ref_num_ops(void* h) {
    RefCtx* c = (RefCtx*) h;
    return (int) c->psg->get_peel_order().size();
}

unsigned int ref_peel_cost(void* h) {
    RefCtx* c = (RefCtx*) h;
    return c->psg->get_peeling_cost();
}

/* info[0..4] = type (peeling.h:17-23), peelnode, cutset size, #previous, #children; arrays sized N */
void ref_op_info(void* h, int i, int* info, int* cutset, int* previous, int* children) {
    RefCtx* c = (RefCtx*) h;
    PeelOperation& op = c->psg->get_peel_order()[i];
    info[0] = (int) op.get_type();
    info[1] = (int) op.get_peelnode();
    info[2] = (int) op.get_cutset_size();
    info[3] = (int) op.get_prev_size();
    info[4] = (int) op.get_children_size();
    for(unsigned j = 0; j < op.get_cutset_size(); ++j) cutset[j] = (int) op.get_cutnode(j);
    for(unsigned j = 0; j < op.get_prev_size(); ++j)   previous[j] = (int) op.get_prevfunctions()[j];
    for(unsigned j = 0; j < op.get_children_size(); ++j) children[j] = (int) op.get_children()[j];
}

/* which: 0 = lod_indices, 1 = matrix_indices[locus], 2 = presum_indices[locus]; returns count
   (copies at most cap entries) */
int ref_op_indices(void* h, int i, int which, int locus, int* buf, int cap) {
    RefCtx* c = (RefCtx*) h;
    PeelOperation& op = c->psg->get_peel_order()[i];
    std::vector<int>* v = (which == 0) ? op.get_lod_indices() :
                          (which == 1) ? op.get_matrix_indices(locus) : op.get_presum_indices(locus);
    int n = (int) v->size();
    for(int j = 0; j < n && j < cap; ++j) buf[j] = (*v)[j];
    return n;
}

/* ---- descent graph --------------------------------------------------------------- */

int ref_dg_random(void* h) {
    RefCtx* c = (RefCtx*) h;
    return c->dg->random_descentgraph() ? 1 : 0;
}

/* int[M][N][2] exactly as DescentGraph::data (descent_graph.h:35-37) */
void ref_dg_get(void* h, int* out) {
    RefCtx* c = (RefCtx*) h;
    memcpy(out, c->dg->get_internal_ptr(), c->dg->get_internal_size());
}

void ref_dg_set(void* h, const int* in) {
    RefCtx* c = (RefCtx*) h;
    memcpy(c->dg->get_internal_ptr(), in, c->dg->get_internal_size());
}

double ref_dg_likelihood(void* h) {
    RefCtx* c = (RefCtx*) h;
    Quiet q(quiet_mode());
    return c->dg->get_likelihood();
}

double ref_dg_recombination_prob(void* h, int locus) {
    RefCtx* c = (RefCtx*) h;
    return c->dg->get_recombination_prob(locus, false);
}

double ref_dg_marker_transmission(void* h) {
    RefCtx* c = (RefCtx*) h;
    return c->dg->get_marker_transmission();
}

void ref_sequential_imputation(void* h, int iterations) {
    RefCtx* c = (RefCtx*) h;
    Quiet q(quiet_mode());
    SequentialImputation si(c->ped, &c->map, c->psg, c->dm.is_sexlinked());
    si.parallel_run(*c->dg, iterations);
}

/* MarkovChain::run on the context's descent graph (mutated in place).  raw[(M-1)*n] gets the
   log-sum accumulators, lodout the normalised LOD (lod_score.h:86-88); returns ln P(T). */
double ref_chain_run(void* h, int burnin, int iterations, int scoring_period, double lsampler_prob,
                     double* raw, double* lodout, int* count) {
    RefCtx* c = (RefCtx*) h;
    Quiet q(quiet_mode());
    struct mcmc_options o = c->opt;
    o.burnin = burnin;
    o.iterations = iterations;
    o.scoring_period = scoring_period;
    o.lsampler_prob = lsampler_prob;
    o.sex_linked = c->dm.is_sexlinked();
    MarkovChain chain(c->ped, &c->map, c->psg, o, 0);
    LODscores* lod = chain.run(*c->dg);
    unsigned n = c->map.get_lodscore_count();
    for(unsigned l = 0; l + 1 < c->map.num_markers(); ++l) {
        for(unsigned k = 0; k < n; ++k) {
            if(raw)    raw[l*n + k] = lod->get_raw(l*n + k);
            if(lodout) lodout[l*n + k] = lod->get(l, k);
        }
    }
    if(count) *count = (int) lod->get_count();
    double tp = lod->trait_prob;
    delete lod;
    return tp;
}

/* ---- L-sampler ------------------------------------------------------------------- */

/* Forward peel at `locus` on the current descent graph.  matrices: concatenated dense pmatrix
   of every op (4^c doubles each, in peel order); presums: concatenated pmatrix_presum
   (4^(c+1) each).  mode 0 = set_locus_minimal (MCMC), 1 = set_locus(l, il, ir) (sequential
   imputation).  Returns the last op's result.  (locus_sampler2.cc:128-136) */
double ref_ls_forward(void* h, int locus, int mode, int ignore_left, int ignore_right,
                      double* matrices, double* presums) {
    RefCtx* c = (RefCtx*) h;
    need_samplers(c);
    LocusSampler* ls = c->ls;
    if(mode == 0) {
        ls->set_locus(locus, false, false);
        ls->set_locus_minimal(locus);
    }
    else ls->set_locus(locus, ignore_left != 0, ignore_right != 0);

    size_t mo = 0, po = 0;
    for(unsigned i = 0; i < ls->rfunctions.size(); ++i) {
        SamplerRfunction& rf = ls->rfunctions[i];
        rf.evaluate(c->dg, 0);
        if(matrices) memcpy(matrices + mo, rf.pmatrix.data, sizeof(double) * rf.pmatrix.size);
        if(presums)  memcpy(presums + po, rf.pmatrix_presum.data, sizeof(double) * rf.pmatrix_presum.size);
        mo += rf.pmatrix.size;
        po += rf.pmatrix_presum.size;
    }
    return ls->rfunctions.back().get_result();
}

/* one reference L-sampler update of `locus` (locus_sampler2.cc:128-181), GSL mt19937 draws */
void ref_ls_step(void* h, int locus) {
    RefCtx* c = (RefCtx*) h;
    need_samplers(c);
    c->ls->set_locus_minimal(locus);
    c->ls->step(*c->dg, locus);
}

/* Backward pass of the reference on an already-forward-peeled locus (call ref_ls_forward
   first): returns the sampled genotype per person in pmk[N] and, per op (peel order), the four
   un-normalised presum values read by SamplerRfunction::sample (sampler_rfunction.cc:159-166). */
void ref_ls_backward_trace(void* h, int* pmk_out, double* dist4) {
    RefCtx* c = (RefCtx*) h;
    LocusSampler* ls = c->ls;
    std::vector<int> pmk(c->ped->num_members(), -1);
    for(int i = (int) ls->rfunctions.size() - 1; i >= 0; --i) {
        SamplerRfunction& rf = ls->rfunctions[i];
        for(int g = 0; g < 4; ++g) {
            pmk[rf.peel_id] = g;
            dist4[i*4 + g] = rf.pmatrix_presum.get(pmk);
        }
        pmk[rf.peel_id] = -1;
        rf.sample(pmk);
    }
    for(unsigned i = 0; i < pmk.size(); ++i) pmk_out[i] = pmk[i];
}

/* LocusSampler::sample_meiosis_indicators (locus_sampler2.cc:93-125) with a caller-supplied
   genotype assignment; writes the context's descent graph at the sampler's current locus */
void ref_ls_sample_indicators(void* h, const int* pmk_in) {
    RefCtx* c = (RefCtx*) h;
    std::vector<int> pmk(pmk_in, pmk_in + c->ped->num_members());
    c->ls->sample_meiosis_indicators(pmk, *c->dg);
}

/* ---- M-sampler and descent-graph likelihood ------------------------------------------ */

/* FounderAlleleGraph4 of one locus on the current descent graph (founder_allele_graph4.cc:548-572,
   :34-424): edge_list[2N] and the likelihood */
double ref_fag(void* h, int locus, int* edge) {
    RefCtx* c = (RefCtx*) h;
    FounderAlleleGraph4 f(c->ped, &c->map, c->dm.is_sexlinked());
    f.set_sequence(&c->dg->seq);
    f.set_locus(locus);
    f.reset(*c->dg);
    for(unsigned i = 0; i < 2 * c->ped->num_members(); ++i) edge[i] = f.edge_list[i];
    return f.likelihood();
}

/* same graph after FounderAlleleGraph4::flip(dg, person, parent) (:574-598) */
double ref_fag_flipped(void* h, int locus, int person, int parent, int* edge) {
    RefCtx* c = (RefCtx*) h;
    FounderAlleleGraph4 f(c->ped, &c->map, c->dm.is_sexlinked());
    f.set_sequence(&c->dg->seq);
    f.set_locus(locus);
    f.reset(*c->dg);
    f.flip(*c->dg, person, static_cast<enum parentage>(parent));
    for(unsigned i = 0; i < 2 * c->ped->num_members(); ++i) edge[i] = f.edge_list[i];
    return f.likelihood();
}

/* MarkovChain::_init's m_ordering (markov_chain.cc:68-80); returns the count */
int ref_ms_ordering(void* h, int* out) {
    RefCtx* c = (RefCtx*) h;
    int n = 0;
    unsigned num_meioses = 2 * (c->ped->num_members() - c->ped->num_founders());
    for(unsigned i = 0; i < num_meioses; ++i) {
        unsigned person_id = c->ped->num_founders() + (i / 2);
        enum parentage p = static_cast<enum parentage>(i % 2);
        if(not c->ped->get_by_index(person_id)->safe_to_ignore_meiosis(p)) out[n++] = (int) i;
    }
    return n;
}

/* MeiosisSampler::reset (meiosis_sampler.cc:17-39) on the context's descent graph */
void ref_ms_reset(void* h, int parameter) {
    RefCtx* c = (RefCtx*) h;
    if(!c->ms) c->ms = new MeiosisSampler(c->ped, &c->map, c->dm.is_sexlinked());
    c->ms->reset(*c->dg, parameter);
}

/* MeiosisSampler::step (meiosis_sampler.cc:104-191).  Before the call the calling thread's
   mt19937 state is cloned and the next `cap` uniforms it would produce are written to us[], so a
   restatement can be fed the very same draws.  Afterwards raw[M][2] / fb[M][2] receive
   raw_matrix / fb_matrix as the step left them. */
void ref_ms_step(void* h, int parameter, double* us, int cap, double* raw, double* fb) {
    RefCtx* c = (RefCtx*) h;
    if(us) {
        gsl_rng clone = *r[get_thread_num()];
        for(int i = 0; i < cap; ++i) us[i] = gsl_rng_uniform(&clone);
    }
    c->ms->step(*c->dg, parameter);
    unsigned M = c->map.num_markers();
    if(raw) for(unsigned i = 0; i < 2 * M; ++i) raw[i] = c->ms->raw_matrix[i];
    if(fb)  for(unsigned i = 0; i < 2 * M; ++i) fb[i] = c->ms->fb_matrix[i];
}

void ref_ms_raw(void* h, double* raw) {
    RefCtx* c = (RefCtx*) h;
    unsigned M = c->map.num_markers();
    for(unsigned i = 0; i < 2 * M; ++i) raw[i] = c->ms->raw_matrix[i];
}

/* `reps` M-sweeps exactly as markov_chain.cc:342-349 runs them; returns seconds */
double ref_bench_msweeps(void* h, int reps) {
    RefCtx* c = (RefCtx*) h;
    Quiet q(quiet_mode());
    std::vector<int> order(2 * c->ped->num_members());
    order.resize(ref_ms_ordering(h, &order[0]));
    MeiosisSampler ms(c->ped, &c->map, c->dm.is_sexlinked());
    double t0 = omp_get_wtime();
    for(int k = 0; k < reps; ++k) {
        random_shuffle(order.begin(), order.end());
        ms.reset(*c->dg, order[0]);
        for(unsigned j = 0; j < order.size(); ++j) ms.step(*c->dg, order[j]);
    }
    return omp_get_wtime() - t0;
}

/* ---- ELOD ------------------------------------------------------------------------------ */

/* Elod(pedfile, options).run() (elod.h:33-107, elod.cc:19-85) with the reference's own generator
   (seeded from /dev/urandom by the constructor, then re-seeded here for reproducibility) */
double ref_elod2(const char* pedfile, double frequency, const double* penetrance, double separation, int replicates,
                 int sex_linked, int affected_only, unsigned int seed);
double ref_elod(const char* pedfile, double frequency, const double* penetrance, double separation, int replicates,
                int sex_linked, unsigned int seed) {
    return ref_elod2(pedfile, frequency, penetrance, separation, replicates, sex_linked, 0, seed);
}

/* the same with --elod -a (options.affected_only, elod.h:85-87) */
double ref_elod2(const char* pedfile, double frequency, const double* penetrance, double separation, int replicates,
                 int sex_linked, int affected_only, unsigned int seed) {
    Quiet q(quiet_mode());
    struct mcmc_options o;
    o.elod = true;
    o.elod_frequency = frequency;
    o.elod_penetrance = std::vector<double>(penetrance, penetrance + 3);
    o.elod_marker_separation = separation;
    o.elod_replicates = replicates;
    o.sex_linked = sex_linked != 0;
    o.affected_only = affected_only != 0;
    o.peelopt_iterations = 20000;
    Elod e(pedfile, o);
    rng_ready = true;
    rng_threads = get_max_threads();
    for(int i = 0; i < rng_threads; ++i) gsl_rng_set(r[i], seed + 7919u * (unsigned) i);
    srand(seed);
    return e.run();
}

/* ---- LOD scoring ----------------------------------------------------------------- */

double ref_calc_trait_prob(void* h) {
    RefCtx* c = (RefCtx*) h;
    need_samplers(c);
    return c->peeler->calc_trait_prob();
}

/* Peeler::process for one interval without the accumulate (peeler.cc:79-103):
   result[k] = last R-function value, prob[k] = log(result) - recomb - transmission.
   matrices (optional): concatenated dense pmatrix of every op for position `dump_k`. */
void ref_lod_interval(void* h, int interval, double* result, double* prob, int dump_k, double* matrices) {
    RefCtx* c = (RefCtx*) h;
    need_samplers(c);
    Peeler* p = c->peeler;
    p->set_locus(interval);
    unsigned n = c->map.get_lodscore_count();
    for(unsigned k = 0; k < n; ++k) {
        for(unsigned j = 0; j < p->rfunctions.size(); ++j) {
            p->rfunctions[j].set_thetas(k + 1);
            p->rfunctions[j].evaluate(c->dg, k + 1);
        }
        double res = p->rfunctions.back().get_result();
        result[k] = res;
        prob[k] = log(res) - c->dg->get_recombination_prob(interval, false) - c->dg->get_marker_transmission();
        if(matrices && (int) k == dump_k) {
            size_t mo = 0;
            for(unsigned j = 0; j < p->rfunctions.size(); ++j) {
                TraitRfunction& rf = p->rfunctions[j];
                memcpy(matrices + mo, rf.pmatrix.data, sizeof(double) * rf.pmatrix.size);
                mo += rf.pmatrix.size;
            }
        }
    }
}

/* ---- CPU baseline timing (markov_chain.cc:209-266, :375-383) ----------------------- */

/* `reps` full L-sweeps over all loci with the reference's own schedulers; returns seconds.
   lgroups < 0 -> run_old_lsampler, else run_scalable_lsampler with that many stride groups. */
double ref_bench_lsweeps(void* h, int reps, int lgroups) {
    RefCtx* c = (RefCtx*) h;
    Quiet q(quiet_mode());
    struct mcmc_options o = c->opt;
    o.burnin = 0; o.iterations = 0; o.sex_linked = c->dm.is_sexlinked();
    MarkovChain chain(c->ped, &c->map, c->psg, o, 0);
    std::vector<int> groups;
    for(int i = 0; i < lgroups; ++i) groups.push_back(i);
    double t0 = omp_get_wtime();
    for(int k = 0; k < reps; ++k) {
        if(lgroups < 0) chain.run_old_lsampler(*c->dg);
        else chain.run_scalable_lsampler(*c->dg, groups, lgroups);
    }
    double t1 = omp_get_wtime();
    delete chain.lod;
    return t1 - t0;
}

/* `reps` scoring passes over all M-1 intervals, the omp loop of markov_chain.cc:375-383 */
double ref_bench_lodpasses(void* h, int reps) {
    RefCtx* c = (RefCtx*) h;
    Quiet q(quiet_mode());
    struct mcmc_options o = c->opt;
    o.burnin = 0; o.iterations = 0; o.sex_linked = c->dm.is_sexlinked();
    MarkovChain chain(c->ped, &c->map, c->psg, o, 0);
    double t0 = omp_get_wtime();
    for(int k = 0; k < reps; ++k) {
        int thread_num = 0;
        #pragma omp parallel private(thread_num)
        {
            thread_num = get_thread_num();
            #pragma omp for
            for(int j = 0; j < int(c->map.num_markers() - 1); ++j) {
                chain.peelers[thread_num]->set_locus(j);
                chain.peelers[thread_num]->process(c->dg);
            }
        }
    }
    double t1 = omp_get_wtime();
    delete chain.lod;
    return t1 - t0;
}

/* The loop of MarkovChain::run (markov_chain.cc:335-390) over the reference's own LocusSampler / MeiosisSampler /
   Peeler objects, for `iterations` iterations: each an L-sweep with probability lsampler_prob (scheduler chosen as
   optimal_num_lgroups would, from `sched_trials` timed sweeps of each kind instead of its 100: lgroups_out reports
   it; sched_trials = -g forces g stride groups, 0 the old sampler), else an M-sweep over the shuffled m_ordering; after every `scoring_period` iterations one scoring pass over
   the first `lod_intervals` intervals (all M-1 if <= 0; a bounded sample keeps a bench step short, the cost per
   interval does not depend on how many are scored).  Returns the seconds of the loop alone (set-up and scheduler
   trials excluded, as the driver's own clock excludes the reference's start-up); counts[0..2] = L-sweeps, M-sweeps,
   scoring passes. */
double ref_bench_chain(void* h, int iterations, int scoring_period, double lsampler_prob, int lod_intervals,
                       int sched_trials, int* lgroups_out, int* counts) {
    RefCtx* c = (RefCtx*) h;
    Quiet q(quiet_mode());
    struct mcmc_options o = c->opt;
    o.burnin = 0; o.iterations = iterations; o.scoring_period = scoring_period; o.lsampler_prob = lsampler_prob;
    o.sex_linked = c->dm.is_sexlinked();
    MarkovChain chain(c->ped, &c->map, c->psg, o, 0);
    DescentGraph& dg = *c->dg;
    int num_lgroups = -1;
    if(sched_trials < 0) num_lgroups = -sched_trials;          /* the caller's earlier choice */
    else if(get_max_threads() > 1 && sched_trials > 0) {
        /* markov_chain.cc:269-311 with fewer trials */
        double best = DBL_MAX;
        for(int g = 2; g <= 10; ++g) {
            if(g == 2) {
                double t0 = omp_get_wtime();
                for(int k = 0; k < sched_trials; ++k) chain.run_old_lsampler(dg);
                best = omp_get_wtime() - t0;
                continue;
            }
            std::vector<int> groups;
            for(int i = 0; i < g; ++i) groups.push_back(i);
            double t0 = omp_get_wtime();
            for(int k = 0; k < sched_trials; ++k) chain.run_scalable_lsampler(dg, groups, g);
            double t = omp_get_wtime() - t0;
            if(t < best) { best = t; num_lgroups = g; }
        }
    }
    if(lgroups_out) *lgroups_out = num_lgroups;
    std::vector<int> lgroups;
    for(int i = 0; i < num_lgroups; ++i) lgroups.push_back(i);
    const int n_int = (lod_intervals > 0 && lod_intervals < int(c->map.num_markers() - 1)) ? lod_intervals : int(c->map.num_markers() - 1);
    int nl = 0, nm = 0, ns = 0;
    double t0 = omp_get_wtime();
    for(int i = 0; i < iterations; ++i) {
        if(get_random() < lsampler_prob) {
            if(num_lgroups == -1) chain.run_old_lsampler(dg);
            else chain.run_scalable_lsampler(dg, lgroups, num_lgroups);
            ++nl;
        }
        else {
            random_shuffle(chain.m_ordering.begin(), chain.m_ordering.end());
            chain.msampler.reset(dg, chain.m_ordering[0]);
            for(unsigned int j = 0; j < chain.m_ordering.size(); ++j) chain.msampler.step(dg, chain.m_ordering[j]);
            ++nm;
        }
        if(((i + 1) % scoring_period) == 0) {
            int thread_num = 0;
            #pragma omp parallel private(thread_num)
            {
                thread_num = get_thread_num();
                #pragma omp for
                for(int j = 0; j < n_int; ++j) {
                    chain.peelers[thread_num]->set_locus(j);
                    chain.peelers[thread_num]->process(&dg);
                }
            }
            ++ns;
        }
    }
    double t1 = omp_get_wtime();
    if(counts) { counts[0] = nl; counts[1] = nm; counts[2] = ns; }
    delete chain.lod;
    return t1 - t0;
}

#ifdef REF_WITH_GPU
/* The reference's own GPU path for LOD scoring (-g): GPULodscores::calculate / get_results (gpu_lodscores.cc:598-637)
   driving lodscore_kernel (cuda_lodscore.cu:389-467), compiled for sm_100a by `make refgpu`.  Runs `reps` scoring
   passes on the context's descent graph; returns the seconds per pass of calculate() + the final block_until_finished()
   (the reference's synchronous pageable copy of the graph included, as in its own loop), writes the normalised LOD
   table (lod_score.h:86-88) of those passes to lodout[(M-1)*n].  setup_s receives the constructor time (device mirror
   of every (op, locus) R-function + its block-size autotune, gpu_lodscores.cc:396-596).
   NOTE: ~GPULodscores calls cudaDeviceReset(): run this in a process of its own. */
double ref_gpu_lod_bench(void* h, int reps, double* lodout, double* setup_s, int* threads_chosen) {
    RefCtx* c = (RefCtx*) h;
    need_samplers(c);
    struct mcmc_options o = c->opt;
    o.sex_linked = c->dm.is_sexlinked();
    o.use_gpu = true;
    const double tp = c->peeler->calc_trait_prob();
    double t0 = omp_get_wtime();
    GPULodscores* g = new GPULodscores(c->ped, &c->map, c->psg, o, tp);
    g->block_until_finished();
    if(setup_s) *setup_s = omp_get_wtime() - t0;
    if(threads_chosen) *threads_chosen = g->num_lodscore_threads;
    g->calculate(*c->dg);                     /* warm-up pass */
    g->block_until_finished();
    t0 = omp_get_wtime();
    for(int k = 0; k < reps; ++k) g->calculate(*c->dg);
    g->block_until_finished();
    const double secs = (omp_get_wtime() - t0) / (reps > 0 ? reps : 1);
    LODscores lod(&c->map);
    g->get_results(&lod);
    unsigned n = c->map.get_lodscore_count();
    if(lodout)
        for(unsigned l = 0; l + 1 < c->map.num_markers(); ++l)
            for(unsigned k = 0; k < n; ++k) lodout[l * n + k] = lod.get(l, k);
    /* the destructor resets the device; leak the object instead and let the process exit clean it up */
    return secs;
}
#endif

} // extern "C"
