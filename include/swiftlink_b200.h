/*
 * swiftlink_b200.h -- C ABI of the B200-native peeling hot path (libswiftlink_b200.so).
 *
 * This is the drop-in boundary for the reference's `-g` path.  The reference crosses from C++
 * into CUDA through the extern "C" block of src/cuda_common.h:232-262, whose wrappers all take a
 * device pointer to a hand-assembled `struct gpu_state` (cuda_common.h:56-83).  Here the same
 * roles are exported with an opaque handle and flat host arrays instead; each entry point
 * names the reference interface it replaces.  No torch / C++ types appear in any signature.
 *
 * Conventions
 *   - every function returns SLK_OK (0) or a negative slk_status; slk_last_error() returns a
 *     message for the calling thread.  (The reference prints and abort()s,
 *     gpu_lodscores.cc:30-37; the C++ classes in swiftlink_b200/csrc/host keep that behaviour
 *     on top of these codes.)
 *   - genotype code: UU=0, AA=1, AU=2, UA=3 (trait.h:21-27, the CPU encoding; the reference
 *     GPU's permuted {AA,BA,AB,BB} order, cuda_common.h:25-30, is not used anywhere).
 *   - descent graph at the boundary: int32[M][N][2], offset = 2N*locus + 2*person + parent,
 *     values 0/1 (descent_graph.h:35-37).  LOD accumulators: natural-log log-sum-exp values,
 *     -DBL_MAX = empty (cuda_common.cu:25).
 *   - all launches are asynchronous on the chain's stream; slk_chain_sync() or any call that
 *     returns data to the host synchronises (as GPULodscores::get_results does,
 *     gpu_lodscores.cc:621-637).
 */
#ifndef SWIFTLINK_B200_H
#define SWIFTLINK_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLK_ABI_VERSION 4

typedef enum {
    SLK_OK = 0,
    SLK_ERR_INVALID = -1,        /* bad argument / malformed plan */
    SLK_ERR_CUDA = -2,           /* CUDA runtime error (message has the cudaError string) */
    SLK_ERR_NO_DEVICE = -3,      /* no usable sm_100 device: the product has no CPU fallback */
    SLK_ERR_ZERO_LIKELIHOOD = -4,/* an L-sampler peel returned 0 (locus_sampler2.cc:137-142) */
    SLK_ERR_NONPOSITIVE_TRAIT = -5, /* a trait peel returned <= 0 (peeler.cc:92-95) */
    SLK_ERR_UNSUPPORTED = -6,
    SLK_ERR_ILLEGAL_GRAPH = -7   /* descent graph with likelihood 0 given to the M-sampler (meiosis_sampler.cc:31-34) */
} slk_status;

enum { SLK_UU = 0, SLK_AA = 1, SLK_AU = 2, SLK_UA = 3 };
enum { SLK_CHILD_PEEL = 1, SLK_PARENT_PEEL = 2, SLK_PARTNER_PEEL = 3, SLK_LAST_PEEL = 4 }; /* peeling.h:17-23 */
enum { SLK_UNSEXED = 0, SLK_MALE = 1, SLK_FEMALE = 2 };                                     /* types.h:25-29 */
enum { SLK_UNTYPED = 0, SLK_HETERO = 1, SLK_HOMOZ_A = 2, SLK_HOMOZ_B = 3 };                 /* genotype.h:26-31 */

#define SLK_MAX_CUTSET   10      /* 4^10 cells; the reference has no explicit bound */
#define SLK_MAX_PREV     8
#define SLK_MAX_CHILDREN 10

/* One elimination step, the fields of PeelOperation (peeling.h:25-39) that the path reads.
   The index tables (assignments / matrix_indices / presum_indices / lod_indices) are NOT part of
   the hand-off: the device derives validity from the elimination masks and disease
   probabilities, which is what lets M = 10k fit ("plan scaling"). */
typedef struct {
    int32_t type;                              /* SLK_*_PEEL */
    int32_t peelnode;
    int32_t ncut;
    int32_t cutset[SLK_MAX_CUTSET];
    int32_t nprev;
    int32_t prev[SLK_MAX_PREV];                /* earlier op ids, in the reference's order */
    int32_t nchild;
    int32_t children[SLK_MAX_CHILDREN];        /* PARENT_PEEL only (peeling.h:153-171) */
} slk_peel_op;

/* Everything the path reads from Pedigree / Person / GeneticMap / DiseaseModel /
   GenotypeElimination / PeelSequenceGenerator, as flat host arrays (copied during the call). */
typedef struct {
    int32_t n_members;                         /* N, founders first (pedigree.cc:154-162) */
    int32_t n_founders;                        /* F */
    int32_t n_markers;                         /* M >= 2 */
    int32_t n_lod;                             /* -n, positions per interval (>= 1) */
    int32_t sex_linked;                        /* -X */
    const int32_t* mother;                     /* [N], -1 for founders */
    const int32_t* father;                     /* [N] */
    const int32_t* sex;                        /* [N] SLK_MALE / SLK_FEMALE / SLK_UNSEXED */
    const int32_t* typed;                      /* [N] Person::istyped() */
    const int32_t* prior_as_founder;           /* [N] Person::isfounder() AS SEEN BY populate_trait_prob_cache.
                                                  In the reference this is 1 for every person: the cache is
                                                  filled while the ped file is parsed (pedigree_parser.cc:153),
                                                  before parent ids are resolved (pedigree.cc:164-193), so
                                                  untyped non-founders get the population genotype prior too.
                                                  NULL means "all 1" (reference behaviour). */
    const uint8_t* genotypes;                  /* [N][M] SLK_UNTYPED.. (Person::get_marker) */
    const double* disease_prob;                /* [N][4] Person::get_disease_prob (person.cc:85-119) */
    const double* marker_prob;                 /* [M][4] Snp::prob   (genetic_map.h:68-74) */
    const double* marker_xprob;                /* [M][4] Snp::x_male_prob (genetic_map.h:76-86) */
    const double* theta;                       /* [M-1] GeneticMap::get_theta */
    const double* partial_theta;               /* [M-1] GeneticMap::get_theta_partial_raw */
    const uint8_t* elimination;                /* [M][N] GenotypeElimination masks AA=8,AB=4,BA=2,BB=1
                                                  (genotype.h:18-24, elimination.cc:393-395) */
    int32_t n_ops;                             /* == N */
    const slk_peel_op* ops;                    /* PeelSequenceGenerator::get_peel_order() */
    const double* minor_freq;                  /* [M] Snp::minor() (genetic_map.h:42); only the M-sampler and the
                                                  descent-graph likelihood read it -- NULL makes those entry
                                                  points fail with SLK_ERR_UNSUPPORTED (ABI 1 callers) */
    int32_t disease_prior_locus_plus1;         /* 0 = none.  l + 1: at marker l every person's genotype prior is their
                                                  disease_prob (Person::copy_disease_probs, person.h:204-208) -- the
                                                  simulated trait locus of the ELOD calculation (elod.h:83) */
    const double* person_prior;                /* [N][4] or NULL (ABI 4).  The genotype prior used at marker
                                                  disease_prior_locus_plus1 - 1: what Person::copy_disease_probs copied, i.e. the
                                                  disease probabilities AT THE TIME OF THE COPY.  The ELOD set-up copies before
                                                  `-a` turns unaffected people into unknowns (elod.h:83-87), so the simulation uses
                                                  the real phenotypes while the trait peel scores them as unknown.  NULL: the
                                                  current disease_prob (identical unless affection changed after the copy). */
} slk_problem;

typedef struct slk_plan slk_plan;              /* device-resident flattened peel plan (shared, read-only) */
typedef struct slk_chain slk_chain;            /* one Markov chain: descent graph, LOD table, RNG key */

/* ---- library ------------------------------------------------------------------------ */

int slk_abi_version(void);
const char* slk_last_error(void);
/* number of usable devices; 0 means every compute call will fail with SLK_ERR_NO_DEVICE */
int slk_device_count(void);

/* ---- plan: replaces GPULodscores::init/gpu_init (gpu_lodscores.cc:106-160, :396-507) and
        GPUMarkovChain's state assembly (gpu_markov_chain.cc:133-700) ---------------------- */

int slk_plan_create(const slk_problem* problem, int device, slk_plan** out);
void slk_plan_destroy(slk_plan* plan);

/* Host-only: flattens the peel sequence exactly as slk_plan_create does and returns the same
   statistics, without touching a device (plan-logic tests run on machines with no GPU). */
int slk_plan_validate(const slk_problem* problem, double* stats, int cap);

/* Plan statistics for roofline accounting.  out[] receives, in order:
   0 n_ops, 1 sum 4^c, 2 sum 4^(c+1), 3 F_L flops per locus update, 4 F_T flops per trait
   position, 5 sampler forward levels, 6 sampler backward levels, 7 trait levels,
   8 sampler arena doubles, 9 trait arena doubles (after liveness reuse), 10 trait valid cells,
   11 max cutset, 12 sampler team threads, 13 trait team threads, 14 sampler arena doubles kept
   in shared memory, 15 trait arena doubles kept in shared memory, 16/17 resident CTAs per SM of
   the sampler / trait kernel, 18/19 their dynamic shared memory per CTA in bytes (16-19 are 0
   from slk_plan_validate).  Returns how many it wrote. */
int slk_plan_stats(const slk_plan* plan, double* out, int cap);

/* ---- chain state ---------------------------------------------------------------------- */

/* seed + chain key the Philox streams: draws depend only on (seed, chain, iteration, locus,
   slot), never on the launch geometry. */
int slk_chain_create(slk_plan* plan, uint64_t seed, uint32_t chain_id, slk_chain** out);
void slk_chain_destroy(slk_chain* chain);
/* run this chain's work on an existing CUDA stream (a cudaStream_t passed as void*); NULL
   restores the chain's own stream */
int slk_chain_set_stream(slk_chain* chain, void* cuda_stream);
int slk_chain_sync(slk_chain* chain);         /* replaces GPULodscores::block_until_finished (:609-619) */

/* replaces the cudaMemcpy of DescentGraph::get_internal_ptr() in GPULodscores::calculate
   (gpu_lodscores.cc:602) / copy_to_gpu / copy_from_gpu */
int slk_dg_upload(slk_chain* chain, const int32_t* dg);
int slk_dg_download(slk_chain* chain, int32_t* dg);
/* exchange the descent graphs of two chains of the same device and dimensions in O(1) (the
   Metropolis-coupled swap of Mc3::run, mc3.cc:155-159, `swap(graphs[rand], graphs[rand+1])`); the
   plans may differ (heated maps).  Both streams are synchronised first. */
int slk_dg_swap(slk_chain* a, slk_chain* b);

/* ---- L-sampler: replaces run_gpu_lsampler_kernel(numblocks, numthreads, state, window_length,
        offset) (cuda_common.h:236, cuda_lsampler.cu:403-449) ------------------------------- */

/* Block-Gibbs update of every locus l with l % window_length == offset (window_length >= 2).
   Equivalent to LocusSampler::set_locus_minimal(l); step(dg, l) for each such l. */
int slk_lsampler_window(slk_chain* chain, uint64_t iteration, int window_length, int offset);
/* One full sweep = both parity classes, first class drawn from the chain's stream
   (gpu_markov_chain.cc:1101-1126 shuffles {0,1} the same way). */
int slk_lsampler_sweep(slk_chain* chain, uint64_t iteration);

/* Independent draw of every locus from its single-locus posterior (neighbours ignored): the
   reference's start state when -s 0, LocusSampler::locus_by_locus (locus_sampler2.cc:183-207,
   sequential_imputation.cc:52-62).  One launch over all loci; needs no previous graph. */
int slk_lsampler_locus_by_locus(slk_chain* chain, uint64_t iteration);

/* One sequential-imputation run: LocusSampler::start_from(dg, start_locus)
   (locus_sampler2.cc:209-241) -- the start locus drawn on its own, then the loci to its left
   each conditioned on its right neighbour, then the loci to its right each conditioned on its
   left neighbour.  Inherently sequential in the loci: one team walks them in one launch.
   `run` takes the place of the iteration in the Philox key.  log_weight (optional) receives the
   importance weight sum ln(likelihood) the reference uses to pick the best of -s runs
   (sequential_imputation.cc:76-108); reading it synchronises. */
int slk_sequential_imputation(slk_chain* chain, uint64_t run, int start_locus, double* log_weight);

/* SequentialImputation::parallel_run (sequential_imputation.cc:47-115): n_runs independent walks of
   LocusSampler::start_from -- run numbers first_run .. first_run + n_runs - 1, walk i starting at start_loci[i] --
   in one launch per wave of resident teams (one team per walk) instead of one SM-filling launch per walk.  Draws,
   weights and graphs are those of n_runs calls of slk_sequential_imputation; log_weights[n_runs] (optional) receives
   every walk's log weight, *best_run the index of the first maximal one, whose graph becomes the chain's graph. */
int slk_sequential_imputation_batch(slk_chain* chain, uint64_t first_run, int n_runs, const int32_t* start_loci,
                                    double* log_weights, int32_t* best_run);

/* ---- M-sampler: replaces run_gpu_msampler_reset_kernel / run_gpu_msampler_likelihood_kernel /
        run_gpu_msampler_sampling_kernel (cuda_common.h:240-244, cuda_msampler2.cu) and the CPU
        MeiosisSampler (meiosis_sampler.cc:17-203) the live reference falls back to --------------- */

/* A meiosis is numbered 2 * (person - F) + parent (0 maternal, 1 paternal), the reference's M-sampler
   `parameter`.  slk_msampler_ordering returns the meioses a sweep visits: MarkovChain::_init's
   m_ordering (markov_chain.cc:68-80), i.e. all but Person::safe_to_ignore_meiosis (person.cc:208-222).
   Writes at most cap entries, returns the count. */
int slk_msampler_ordering(const slk_plan* plan, int32_t* out, int cap);
/* MeiosisSampler::reset: founder-allele-graph likelihood of the chain's current graph at every
   locus, kept on the device and carried from step to step.  Needed after anything but an M-sampler
   step changed the graph. */
int slk_msampler_reset(slk_chain* chain);
/* MeiosisSampler::step(dg, meiosis): whole-chromosome Gibbs update of one meiosis indicator.  The
   indicator draw of locus l is uniform(seed, chain, iteration, l, SLK_SLOT_MEIOSIS + meiosis). */
int slk_msampler_step(slk_chain* chain, uint64_t iteration, int meiosis);
/* One M-sweep (markov_chain.cc:342-349): Philox Fisher-Yates shuffle of the ordering, reset, one step
   per meiosis -- every launch asynchronous on the chain's stream.  Steps are issued in pairs: the
   second meiosis's likelihoods are evaluated under both outcomes of the first, which changes the
   launch count, not the result. */
int slk_msampler_sweep(slk_chain* chain, uint64_t iteration);
/* DescentGraph::get_likelihood (descent_graph.cc:150-156): ln of transmission x recombination x
   founder-allele-graph prior; -DBL_MAX (LOG_ILLEGAL) if some locus has likelihood 0.  Synchronises. */
int slk_dg_likelihood(slk_chain* chain, double* log_likelihood);
/* which kind of sweep iteration `iteration` of this chain is: 1 = L-sweep, 0 = M-sweep
   (get_random() < lsampler_prob, markov_chain.cc:332, drawn from the Philox stream) */
int slk_sweep_is_lsampler(const slk_chain* chain, uint64_t iteration, double lsampler_prob);

/* ---- LOD scoring: replaces run_gpu_lodscoreinit_kernel / run_gpu_lodscore_kernel /
        run_gpu_lodscorenormalise_kernel (cuda_common.h:249-252, cuda_lodscore.cu:389-541) ---- */

int slk_lodscore_init(slk_chain* chain);                       /* all accumulators := -DBL_MAX, count := 0 */
/* score the chain's current descent graph at every (interval, position) and log-sum the
   result into the accumulators: Peeler::process over all intervals (peeler.cc:79-103) */
int slk_lodscore_accumulate(slk_chain* chain);
/* raw accumulators [(M-1)*n_lod] and the number of scoring passes (LODscores::get_raw / count) */
int slk_lodscore_read(slk_chain* chain, double* raw, int32_t* count);
/* device-side (score - ln count - trait_prob) / ln 10 (lod_score.h:86-88) into out[(M-1)*n_lod] */
int slk_lodscore_normalise(slk_chain* chain, double trait_prob, double* out);
/* ln P(trait) without marker data: Peeler::calc_trait_prob (peeler.cc:65-73) */
int slk_trait_likelihood(slk_plan* plan, double* log_prob);

/* ---- ELOD: replaces the replicate loop of Elod::run (elod.cc:55-66) ------------------------------- */

/* `replicates` independent simulations on one device: LocusSampler::start_from(dg1, 1) on the
   sampler plan (three loci: marker, simulated trait locus, marker; all genotypes untyped), the two
   marker rows copied into a two-locus graph (DescentGraph::copy_locus, elod.cc:58-59) and
   Peeler::process on the trait plan (two loci, n_lod = 1), log-summed.  Both plans must describe the
   same pedigree on the same device.  Replicate r draws from the Philox stream (seed, chain, iteration =
   0, locus = 3 r + j).  log_sum receives the LODscores accumulator (natural log, LOG_ZERO if empty),
   count the number of replicates added; LOD = (log_sum - ln count - ln P(T)) / ln 10 (lod_score.h:86-88).
   prob_out (optional, [replicates]) receives every replicate's ln-probability for parity tests. */
int slk_elod_run(slk_plan* sampler_plan, slk_plan* trait_plan, uint64_t seed, uint32_t chain_id,
                 int64_t replicates, double* log_sum, int64_t* count, double* prob_out);
/* the three-locus graphs of replicates [first, first + n) as the run above samples them:
   dg int32[n][3][N][2] (parity tests) */
int slk_debug_elod_graphs(slk_plan* sampler_plan, uint64_t seed, uint32_t chain_id, int64_t first, int32_t n,
                          int32_t* dg);

/* ---- parity hooks (used by tests/; not needed by a caller) ------------------------------- */

/* forward peel of one locus on the chain's descent graph: dense peel matrices (sum 4^c) and
   presum matrices (sum 4^(c+1)) of every op, invalid cells 0; ignore_left/right as in
   SamplerRfunction::set_locus (sampler_rfunction.h:84-112) */
int slk_debug_lsampler_forward(slk_chain* chain, int locus, int ignore_left, int ignore_right,
                               double* matrices, double* presums, double* result);
/* one L-sampler update of one locus; also returns the sampled genotypes pmk[N] and the
   un-normalised sampling vectors dist4[n_ops][4] */
int slk_debug_lsampler_step(slk_chain* chain, uint64_t iteration, int locus, int ignore_left,
                            int ignore_right, int32_t* pmk, double* dist4, double* result);
/* trait peel of one interval: result[n_lod], prob[n_lod] (= ln result - recomb - transmission)
   and, if matrices != NULL, the dense peel matrices of position dump_k */
int slk_debug_lod_interval(slk_chain* chain, int interval, double* result, double* prob,
                           int dump_k, double* matrices);
/* founder allele graph of every locus on the chain's graph, optionally with one meiosis flipped
   (meiosis < 0: none): ln likelihood [M] (-inf where the reference returns 0) and, if edges != NULL,
   the edge lists [M][2N] (FounderAlleleGraph4::reset / flip / likelihood) */
int slk_debug_fag(slk_chain* chain, int meiosis, double* ln_lik, int32_t* edges);
/* after slk_msampler_step: the forward matrix fb[M][2] (meiosis_sampler.cc:134-153, before the
   backward pass) and the carried per-locus ln likelihood [M] of the graph the step left */
int slk_debug_msampler_state(slk_chain* chain, double* fb, double* ln_cur);
/* tuning aid: one three-hypothesis likelihood launch for the pair (meiosis0, meiosis1) that records
   clock64() stamps of lane 0 of every 64th CTA: stamps[0..11][8] = start, tables built, genotypes staged,
   labels done, typed labels kept, graph walked, result written (only the first 12 sampled CTAs are kept);
   followed by the two-step chain kernel of the same pair: stamps[12..14] = CTA 0, stamps[15..17] = last CTA
   (start, then per step: staged, chunk product, CTA scan, cluster sync, recurrence, map scans, applied, synced).
   stamps must hold 160 values. */
int slk_debug_msampler_trace(slk_chain* chain, int meiosis0, int meiosis1, long long* stamps);
/* measurement aid for bench.py: enqueue `reps` launches of one M-sampler kernel for the pair
   (meiosis0, meiosis1): which = 0 the three-hypothesis incremental likelihood kernel, 1 the two-step chain kernel,
   2 the full likelihood kernel (every label recomputed; what a sweep's reset launches once) */
int slk_debug_msampler_launch(slk_chain* chain, int meiosis0, int meiosis1, int which, int reps);
/* tuning aid: one M-sweep (as slk_msampler_sweep) whose launches record %globaltimer stamps.  With n = the number of
   meioses of a sweep: out[8 j + 0..4] = likelihood launch of the pair at order positions (j, j + 1), CTA 0: start, walk
   begins, before the wait for its predecessor, after it, end; out[8 (j + 1) + 0..3] = that pair's chain launch: start,
   before the wait, after it, end; from out[8 (n + 2)] on, for every CTA b of the likelihood launch at order position
   cta_pair: out[.. + 2 b] = start, out[.. + 2 b + 1] = end << 10 | SM id.  cap >= 8 (n + 2) + 8 ceil(M / 32) words. */
int slk_debug_msampler_timeline(slk_chain* chain, uint64_t iteration, int cta_pair, unsigned long long* out, int cap);
/* tuning aid: one production L-sampler window launch (window 2, given offset) that also records
   clock64() stamps of the first team's first locus: start, after staging, after every forward
   level, after every backward level, after the indicators.  Returns the number of stamps. */
int slk_debug_lsampler_trace(slk_chain* chain, uint64_t iteration, int offset, long long* stamps, int cap);
/* Philox block and the uniform draw used by the kernels, evaluated on the device */
int slk_debug_philox(int device, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
int slk_debug_uniform(int device, uint64_t seed, uint32_t chain, uint64_t iteration, uint32_t locus,
                      uint32_t slot, double* out);

/* measured FP64 FMA throughput of the device in TFLOP/s (2 flops per DFMA), for the roofline */
int slk_measure_fp64_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
