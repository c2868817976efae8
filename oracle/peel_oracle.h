/*
 * peel_oracle.h -- CPU restatement of the reference's peeling hot path.
 * TEST INFRASTRUCTURE ONLY: nothing under swiftlink_b200/ may include, link or call this.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function here against
 * the unmodified reference compiled into oracle/_ref/libswiftref.so (all peel matrices, presum
 * matrices, sampling 4-vectors, per-position trait likelihoods, P(T)) on the reference's three
 * example pedigrees, and tests/golden/*.npz freezes a sample of those outputs so the check also
 * runs where oracle/_ref is absent.
 *
 * Inputs are the reference's own shapes (PeelOperation fields, DescentGraph int[M][N][2],
 * GenotypeElimination masks, Person tables) as flat arrays -- deliberately NOT the product's
 * flattened device plan, so that the plan flattening is itself under test.
 */
#ifndef SLK_PEEL_ORACLE_H
#define SLK_PEEL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_UU = 0, ORC_AA = 1, ORC_AU = 2, ORC_UA = 3 };              /* trait.h:21-27 */
enum { ORC_CHILD_PEEL = 1, ORC_PARENT_PEEL = 2, ORC_PARTNER_PEEL = 3, ORC_LAST_PEEL = 4 }; /* peeling.h:17-23 */
enum { ORC_UNSEXED = 0, ORC_MALE = 1, ORC_FEMALE = 2 };              /* types.h:25-29 */
enum { ORC_UNTYPED = 0, ORC_HETERO = 1, ORC_HOMOZ_A = 2, ORC_HOMOZ_B = 3 }; /* genotype.h:26-31 */

#define ORC_MAXC 12

typedef struct {
    int type;
    int peelnode;
    int ncut;   int cutset[ORC_MAXC];
    int nprev;  int prev[ORC_MAXC];
    int nchild; int children[ORC_MAXC];
} orc_op;

typedef struct {
    int N, F, M, nlod, sex_linked;
    const int* mother;            /* [N], -1 for founders */
    const int* father;            /* [N] */
    const int* sex;               /* [N] */
    const double* disease_prob;   /* [N][4]   Person::get_disease_prob            */
    const double* marker_prob;    /* [N][M][4] Person::get_trait_probability       */
    const int* elim;              /* [M][N]   GenotypeElimination masks (AA=8,AB=4,BA=2,BB=1) */
    const double* theta;          /* [M-1]    GeneticMap::get_theta               */
    const double* partial;        /* [M-1]    GeneticMap::get_theta_partial_raw   */
    int nops;
    const orc_op* ops;
    /* M-sampler / founder allele graph inputs (msampler_oracle.c) */
    const int* typed;             /* [N]      Person::istyped()                    */
    const int* genotypes;         /* [N][M]   Person::get_marker (ORC_UNTYPED..)   */
    const double* minor;          /* [M]      Snp::minor(); major = 1 - minor      */
} orc_problem;

/* sizes of the concatenated dense matrices (sum 4^c) and presum matrices (sum 4^(c+1)) */
long orc_matrix_doubles(const orc_problem* p);
long orc_presum_doubles(const orc_problem* p);

/* genotype.cc:91-102 + elimination.cc:393-395 */
int orc_is_legal(const orc_problem* p, int person, int locus, int value);

/* peel_sequence_generator.cc:84-186: index lists of one op; which = 0 lod, 1 matrix, 2 presum.
   Returns the count, writes at most cap entries. */
int orc_op_indices(const orc_problem* p, int op, int which, int locus, int* out, int cap);

/* person.cc:224-299: normalised marker-genotype prior of one person at one locus */
void orc_marker_prob(int isfounder, int typed, int genotype, int xmale, const double mapprob[4], double out[4]);

/* locus_sampler2.cc:128-136 forward half; returns the last op's result */
double orc_ls_forward(const orc_problem* p, const int* dg, int locus, int ignore_left, int ignore_right,
                      double* matrices, double* presums);

/* Whole LocusSampler::step (locus_sampler2.cc:128-159) with Philox draws (philox.h schedule).
   dg is updated in place at `locus`.  pmk_out[N] (optional) = sampled genotypes,
   dist4_out[nops][4] (optional) = un-normalised sampling vectors.  Returns the likelihood. */
double orc_ls_step(const orc_problem* p, int* dg, int locus, int ignore_left, int ignore_right,
                   uint64_t seed, uint32_t chain, uint64_t iteration, int* pmk_out, double* dist4_out);

/* One full L-sweep in the product's schedule: one parity class of loci, then the other, the
   first class chosen by the ORC_SLOT_PHASE draw.  Returns 0, or 1+locus if a likelihood was 0. */
int orc_ls_sweep(const orc_problem* p, int* dg, uint64_t seed, uint32_t chain, uint64_t iteration);

/* LocusSampler::start_from (locus_sampler2.cc:209-241): one sequential-imputation run from
   start_locus, Philox keyed with `run` in place of the iteration.  Returns sum ln(likelihood),
   or -DBL_MAX if a locus had likelihood 0. */
double orc_si_start_from(const orc_problem* p, int* dg, int start_locus, uint64_t seed, uint32_t chain, uint64_t run);

/* locus_sampler2.cc:44-65: P(indicator = 0) for a homozygous parent */
double orc_homo_p0(const orc_problem* p, const int* dg, int locus, int person, int parent,
                   int ignore_left, int ignore_right);

/* peeler.cc:79-103 without the accumulate; result[nlod], prob[nlod]; optional dense matrices of
   position dump_k */
void orc_lod_interval(const orc_problem* p, const int* dg, int interval, double* result, double* prob,
                      int dump_k, double* matrices);

/* peeler.cc:65-73: ln P(trait) with no marker data */
double orc_trait_prob(const orc_problem* p);

/* descent_graph.cc:212-242 and :22,35 */
double orc_recombination_prob(const orc_problem* p, const int* dg, int locus);
double orc_marker_transmission(const orc_problem* p);

/* logarithms.cc:14-29, lod_score.h:74-88 */
double orc_log_sum(double a, double b);
void orc_lod_add(double* scores, int n, const double* prob, int first);
double orc_lod_normalise(double score, int count, double trait_prob);

/* one scoring pass over all intervals (markov_chain.cc:375-383) accumulated into scores[(M-1)*nlod] */
void orc_lod_pass(const orc_problem* p, const int* dg, double* scores, int first);

/* one ELOD replicate (elod.cc:55-61) on a three-locus sampler problem and a two-locus trait problem;
   Philox keyed as if the replicate's loci were rows 3r..3r+2 of one long graph */
double orc_elod_replicate(const orc_problem* p1, const orc_problem* p2, long replicate, uint64_t seed, uint32_t chain, int* dg3);

/* ---- M-sampler and descent-graph likelihood (msampler_oracle.c) ------------------------- */

typedef struct orc_msampler orc_msampler;

void orc_fag_sequence(const orc_problem* p, int* seq);                                  /* meiosis_sampler.cc:41-72 */
void orc_fag_reset(const orc_problem* p, const int* dg, int locus, int* edge);          /* founder_allele_graph4.cc:548-572 */
void orc_fag_flip(const orc_problem* p, const int* dg, int locus, int person, int parent, int* edge); /* :574-598 */
double orc_fag_likelihood(const orc_problem* p, int locus, const int* edge);            /* :34-424 */

orc_msampler* orc_ms_create(const orc_problem* p);
void orc_ms_destroy(orc_msampler* m);
/* MeiosisSampler::reset (meiosis_sampler.cc:17-39); 0, or 1 + locus of an illegal graph */
int orc_ms_reset(orc_msampler* m, const int* dg, int parameter);
/* MeiosisSampler::step (meiosis_sampler.cc:104-191), dg updated in place.  _stream takes the
   caller's uniforms in the reference's order of consumption; the other the Philox schedule. */
int orc_ms_step_stream(orc_msampler* m, int* dg, int parameter, const double* us, int n, int* used);
int orc_ms_step(orc_msampler* m, int* dg, int parameter, uint64_t seed, uint32_t chain, uint64_t iteration);
/* raw_matrix [M][2], fb_matrix after the forward pass [M][2], fb_matrix as the step leaves it
   [M][2], edge lists [M][2N]; any pointer may be NULL */
void orc_ms_state(const orc_msampler* m, double* raw, double* fwd, double* fb, int* edges);
/* markov_chain.cc:68-80: meioses visited by a sweep (pedigree order); returns the count */
int orc_ms_ordering(const orc_problem* p, int* out);
void orc_ms_shuffle(int* v, int n, uint64_t seed, uint32_t chain, uint64_t iteration);
/* one M-sweep (markov_chain.cc:342-349): shuffle, reset, step every meiosis */
int orc_ms_sweep(const orc_problem* p, int* dg, uint64_t seed, uint32_t chain, uint64_t iteration);

/* descent_graph.cc:150-265 */
double orc_dg_sum_prior_prob(const orc_problem* p, const int* dg);
double orc_dg_likelihood(const orc_problem* p, const int* dg);

/* philox.h, exported for the known-answer test */
void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_uniform_draw(uint64_t seed, uint32_t chain, uint64_t iteration, uint32_t locus, uint32_t slot);

#ifdef __cplusplus
}
#endif
#endif
