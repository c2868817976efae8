/*
 * swiftlink_b200_host.h -- C shims over the C++ host side (swiftlink_b200/csrc/host): LINKAGE
 * ped/map/dat parsing, pedigree / map / disease-model tables, genotype elimination and the
 * peel-sequence generator, exposed so that non-C++ callers (the Python tests and bench.py) can
 * build an slk_problem without the oracle.  These mirror, call for call, what
 * LinkageProgram::run_pedigree does before it constructs the chain
 * (linkage_program.cc:110-171: parse -> DescentGraph -> PeelSequenceGenerator).
 *
 * Errors follow the reference: parse problems are printed to stderr and reported by a NULL /
 * zero return; internal inconsistencies abort().
 */
#ifndef SWIFTLINK_B200_HOST_H
#define SWIFTLINK_B200_HOST_H

#include "swiftlink_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct slk_host slk_host;

/* Program::read_and_check_input (program.cc:18-63); exactly one pedigree is expected */
slk_host* slk_host_open(const char* pedfile, const char* mapfile, const char* datfile, int force_sex_linked, int n_lod);
void slk_host_close(slk_host* h);

/* out[0..4] = N, F, M, n_lod, sex_linked */
void slk_host_dims(const slk_host* h, int32_t* out);
/* mother/father (-1 = founder), sex, affection, typed: [N]; disease_prob: [N][4] */
void slk_host_person_table(const slk_host* h, int32_t* mother, int32_t* father, int32_t* sex, int32_t* affection,
                           int32_t* typed, double* disease_prob);
int slk_host_person_name(const slk_host* h, int i, char* buf, int cap);
int slk_host_marker_name(const slk_host* h, int l, char* buf, int cap);
void slk_host_genotypes(const slk_host* h, int32_t* out);                /* [N][M] */
void slk_host_marker_trait_prob(const slk_host* h, double* out);          /* [N][M][4] */
/* gdist [M], minor [M], prob [M][4], xprob [M][4], theta [M-1], partial [M-1] */
void slk_host_map_table(const slk_host* h, double* gdist, double* minor, double* prob, double* xprob,
                        double* theta, double* partial);
void slk_host_disease_model(const slk_host* h, double* out);             /* freq, penetrance[3] */
void slk_host_elim_masks(slk_host* h, int32_t* out);                      /* [M][N] */

/* PeelSequenceGenerator::build_peel_sequence (search) / import of a known elimination order */
void slk_host_build_peel(slk_host* h, int iterations, uint64_t seed);
int slk_host_set_peel(slk_host* h, const uint32_t* seq);                  /* 1 = accepted */
int slk_host_num_ops(const slk_host* h);
uint32_t slk_host_peel_cost(const slk_host* h);
/* info[0..4] = type, peelnode, #cutset, #previous, #children */
void slk_host_op_info(const slk_host* h, int i, int32_t* info, int32_t* cutset, int32_t* previous, int32_t* children);

/* GenotypeElimination::random_descentgraph into int32[M][N][2]; 1 on success */
int slk_host_random_descentgraph(slk_host* h, uint64_t seed, int32_t* dg);

/* the flattened problem (valid until the next build/set_peel or close) */
const slk_problem* slk_host_problem(slk_host* h);

/* LinkageWriter::write for one pedigree: lod[(M-1)*n_lod] already normalised */
int slk_host_write_results(slk_host* h, const char* filename, const double* lod);

/* GPUMarkovChain::run on the device: burnin + iterations sweeps, each an L-sweep with probability
   lsampler_prob and an M-sweep otherwise (markov_chain.cc:332-349), scoring every
   scoring_period-th iteration after burn-in (markov_chain.cc:357-361), dg int32[M][N][2] updated
   in place; lod_out[(M-1)*n_lod] receives the normalised LOD scores.  Returns 0 or an slk_status. */
int slk_host_run_chain(slk_host* h, int device, uint64_t seed, uint32_t chain_id, int burnin, int iterations,
                       int scoring_period, double lsampler_prob, int32_t* dg, double* lod_out, double* trait_prob_out);

/* the -R replicate loop of LinkageProgram::run_pedigree (linkage_program.cc:96-108): `runs` independent chains
   (chain ids 0..runs-1, each started from its own sequential-imputation state), up to `in_flight` of them resident
   on the device at once, each on its own stream; lod_out[(M-1)*n_lod] receives the normalised LOD scores of the
   merged table (LODscores::merge_results in replicate order).  The result does not depend on in_flight. */
int slk_host_run_replicates(slk_host* h, int device, uint64_t seed, int runs, int in_flight, int burnin, int iterations,
                            int scoring_period, double lsampler_prob, int si_iterations, double* lod_out);

/* Mc3::run (mc3.cc:81-200, see swiftlink_b200/csrc/host/mc3.cc for the specified behaviour): a ladder
   of n_chains heated chains on one device, temperatures NULL = the reference's ladder
   1 / (1 + 0.001 * 2^i); swap_success / swap_failure [n_chains] receive the exchange counts of each
   adjacent pair; lod_out the cold chain's normalised LOD scores. */
int slk_host_run_mc3(slk_host* h, int device, uint64_t seed, uint32_t chain_id, int n_chains, int exchange_period,
                     const double* temperatures, int burnin, int iterations, int scoring_period, double lsampler_prob,
                     int si_iterations, double* lod_out, int32_t* swap_success, int32_t* swap_failure);
double slk_host_mc3_temperature(int chain_index, int n_chains, const double* temperatures);

/* One device's share of a `-R` job (LinkageProgram::run_pedigree's replicate loop, linkage_program.cc:96-108): the
   replicates `replicate_ids` (their chain ids; a job of R replicates over W processes places replicate r on process
   r mod W), each a plain chain (mc3_chains <= 1) or a Metropolis-coupled ladder of mc3_chains chains (mc3.cc:81-200),
   all resident on `device` at once.  create() builds the chains and their sequential-imputation start states;
   advance() runs the next n iterations of every replicate (ladders: whole spurts of the exchange period) and returns
   how many were run; results() merges the replicates' RAW log-sum accumulators and counts as LODscores::merge_results
   does (lod_score.h:98-105) -- raw[(M-1)*n_lod], LOG_ZERO (-DBL_MAX) where nothing was scored -- so that tables of
   different processes can be merged the same way (max / sum-exp / sum all-reduces, swiftlink_b200/dist.py), and sums
   the ladders' swap counters into swap_success / swap_failure [n_swap]. */
typedef struct slk_host_job slk_host_job;
slk_host_job* slk_host_job_create(slk_host* h, int device, uint64_t seed, const int32_t* replicate_ids, int n_replicates,
                                  int mc3_chains, int exchange_period, const double* temperatures, int burnin, int iterations,
                                  int scoring_period, double lsampler_prob, int si_iterations);
int slk_host_job_advance(slk_host_job* job, int n_iterations);
int slk_host_job_results(slk_host_job* job, double* raw, int32_t* count, double* trait_prob, int32_t* swap_success,
                         int32_t* swap_failure, int n_swap);
void slk_host_job_destroy(slk_host_job* job);
/* Elod(pedfile, options).run() (elod.h:33-107, elod.cc:19-85): expected LOD of every pedigree of the file by
   simulation on the device; per_pedigree[cap] receives the individual values, the total is returned */
double slk_host_elod(const char* pedfile, double frequency, const double* penetrance, double separation, int replicates,
                     int sex_linked, int affected_only, int peel_iterations, uint64_t seed, int device,
                     double* per_pedigree, int cap);

#ifdef __cplusplus
}
#endif
#endif
