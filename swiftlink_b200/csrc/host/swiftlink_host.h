// swiftlink_host.h -- C++ host side above the C ABI: the classes a SwiftLink caller already
// knows (Pedigree / Person / GeneticMap / DiseaseModel / parsers / GenotypeElimination /
// PeelSequenceGenerator / DescentGraph / LODscores / GPULodscores / GPUMarkovChain), with the
// same names, argument meaning and error behaviour (print to stderr, abort()/exit()) as the
// reference, re-written around flat tables that feed slk_problem directly.
//
// Every numeric table produced here is checked bit-for-bit against the compiled reference in
// tests/test_host.py (person order, disease probabilities, marker priors, thetas,
// elimination masks, peel operations for a given elimination order).
#ifndef SWIFTLINK_HOST_H
#define SWIFTLINK_HOST_H

#include <stdint.h>
#include <cstdio>
#include <string>
#include <vector>

#include "swiftlink_b200.h"

namespace swiftlink {

// ---- enums (types.h:19-35, trait.h:9-27, genotype.h:18-31, peeling.h:17-23) --------------
enum parentage { MATERNAL = 0, PATERNAL = 1, NONE = 2 };
enum sex { UNSEXED = 0, MALE = 1, FEMALE = 2 };
enum affection { UNKNOWN_AFFECTION = 0, UNAFFECTED = 1, AFFECTED = 2 };
enum unphased_trait { TRAIT_HOMO_U = 0, TRAIT_HETERO = 1, TRAIT_HOMO_A = 2 };
enum phased_trait { TRAIT_UU = 0, TRAIT_AA = 1, TRAIT_AU = 2, TRAIT_UA = 3 };
enum unphased_genotype { UNTYPED = 0, HETERO = 1, HOMOZ_A = 2, HOMOZ_B = 3 };
enum peeloperation { NULL_PEEL = 0, CHILD_PEEL = 1, PARENT_PEEL = 2, PARTNER_PEEL = 3, LAST_PEEL = 4 };
enum { GENO_AA = 8, GENO_AB = 4, GENO_BA = 2, GENO_BB = 1 };

const unsigned int UNKNOWN_PARENT = ~0u;

// types.h:55-144 (the fields this code base reads)
struct mcmc_options {
    bool verbose;
    int burnin, iterations, si_iterations, scoring_period, mcmc_runs;
    bool coda_logging;
    std::string coda_prefix;
    int lodscores, peelopt_iterations;
    double lsampler_prob;
    int thread_count;
    bool use_gpu;
    std::string random_filename;
    bool affected_only, sex_linked;
    uint64_t seed;              // Philox seed (replaces the per-thread mt19937 seed file)
    int device;                 // CUDA device of this chain group
    // ELOD (types.h:87-92, defaults.h:4-7)
    bool elod;
    double elod_frequency, elod_penetrance[3], elod_marker_separation;
    int elod_replicates;
    // Metropolis-coupled MCMC (types.h:94-98)
    bool mc3;
    int mc3_number_of_chains, mc3_exchange_period;
    std::vector<double> mc3_temperatures;
    std::string exchange_filename;
    mcmc_options();
};

// ---- disease model (disease_model.h/.cc) ---------------------------------------------------
class DiseaseModel {
    double frequency;
    double penetrance[3];
    bool sexlinked;
 public:
    DiseaseModel() : frequency(0.001), sexlinked(false) { penetrance[0] = penetrance[1] = penetrance[2] = 0.0; }
    DiseaseModel(double freq, const double pen[3], bool sexlink) : frequency(freq), sexlinked(sexlink) {   // disease_model.h:31-44
        for(int i = 0; i < 3; ++i) penetrance[i] = pen[i];
    }
    void set_freq(double f) { frequency = f; }
    void set_penetrance(double p, enum unphased_trait t) { penetrance[t] = p; }
    void set_sexlinked(bool s) { sexlinked = s; }
    double get_freq() const { return frequency; }
    double get_penetrance(enum unphased_trait t) const { return penetrance[t]; }
    bool is_sexlinked() const { return sexlinked; }
    void finish_init() {}
    double get_penetrance_prob2(enum affection a, enum unphased_trait t, enum sex s) const;
    double get_apriori_prob2(enum affection a, enum unphased_trait t, enum sex s) const;
};

// ---- genetic map (genetic_map.h/.cc) -------------------------------------------------------
class Snp {
    std::string name;
    double genetic_distance;
    double major_freq, minor_freq;
    bool maf_set;
    double prob[4], x_male_prob[4];
 public:
    Snp(const std::string& name, double genetic) :
        name(name), genetic_distance(genetic), major_freq(1.0), minor_freq(0.0), maf_set(false) {
        for(int i = 0; i < 4; ++i) prob[i] = x_male_prob[i] = 0.0;
    }
    double major() const { return major_freq; }
    double minor() const { return minor_freq; }
    void set_minor_freq(double m) { minor_freq = m; major_freq = 1.0 - m; maf_set = true; }
    std::string get_name() const { return name; }
    double get_g_distance() const { return genetic_distance; }
    bool is_maf_set() const { return maf_set; }
    void init_probs();
    double get_prob(enum phased_trait pt, bool x_male) const { return x_male ? x_male_prob[pt] : prob[pt]; }
};

class GeneticMap {
    std::vector<Snp> map;
    std::vector<double> thetas, inversethetas, partial_thetas;
    double temperature;
    unsigned int partial_theta_count;
 public:
    explicit GeneticMap(unsigned int partial_theta_count) : temperature(1.0), partial_theta_count(partial_theta_count) {}
    Snp& operator[](int i) { return map[i]; }
    const Snp& operator[](int i) const { return map[i]; }
    double haldane(double m) const;
    double inverse_haldane(double r) const;
    void add(const Snp& s) { map.push_back(s); }
    void add_theta(double d) { thetas.push_back(d); inversethetas.push_back(1.0 - d); }
    std::string get_name(unsigned int i) const { return map[i].get_name(); }
    double get_minor(unsigned int i) const { return map[i].minor(); }
    double get_major(unsigned int i) const { return map[i].major(); }
    double get_prob(unsigned int i, enum phased_trait pt, bool x_male) const { return map[i].get_prob(pt, x_male); }
    unsigned int get_lodscore_count() const { return partial_theta_count; }
    double get_theta(unsigned int i) const { return thetas[i]; }
    double get_inversetheta(unsigned int i) const { return inversethetas[i]; }
    double get_theta_log(unsigned int i) const;
    double get_inversetheta_log(unsigned int i) const;
    double get_genetic_position(unsigned int index, unsigned int offset) const;
    double get_theta_partial(unsigned int index, unsigned int offset) const { return partial_thetas[index] * offset; }
    double get_theta_partial_raw(unsigned int index) const { return partial_thetas[index]; }
    unsigned int num_markers() const { return (unsigned int) map.size(); }
    unsigned int num_thetas() const { return (unsigned int) thetas.size(); }
    bool sanity_check();
    void set_temperature(double t);
};

// ---- pedigree (person.h/.cc, pedigree.h/.cc) -----------------------------------------------
class Pedigree;

class Person {
    std::string id, mother, father;
    enum sex gender;
    enum affection affection_status;
    unsigned int internal_id, maternal_id, paternal_id;
    bool typed;
    double disease_prob[4];
    std::vector<unsigned char> genotypes;            // unphased_genotype per marker
    std::vector<double> genotypes_prob;              // [M][4]
    std::vector<unsigned int> children, mates;       // internal ids
    friend class Pedigree;
 public:
    Person(const std::string& name, const std::string& father_name, const std::string& mother_name,
           enum sex s, enum affection a, const DiseaseModel& dm);
    void init_probs(const DiseaseModel& dm);
    std::string get_id() const { return id; }
    std::string get_mother() const { return mother; }
    std::string get_father() const { return father; }
    unsigned int get_internalid() const { return internal_id; }
    unsigned int get_maternalid() const { return maternal_id; }
    unsigned int get_paternalid() const { return paternal_id; }
    enum sex get_sex() const { return gender; }
    enum affection get_affection() const { return affection_status; }
    enum unphased_genotype get_genotype(unsigned int i) const { return typed ? (enum unphased_genotype) genotypes[i] : UNTYPED; }
    enum unphased_genotype get_marker(unsigned int i) const { return (enum unphased_genotype) genotypes[i]; }
    unsigned int num_markers() const { return (unsigned int) genotypes.size(); }
    unsigned int num_children() const { return (unsigned int) children.size(); }
    unsigned int num_mates() const { return (unsigned int) mates.size(); }
    unsigned int get_child(unsigned int i) const { return children[i]; }
    unsigned int get_mate(unsigned int i) const { return mates[i]; }
    bool is_offspring(unsigned int node) const;
    void add_genotype(enum unphased_genotype g) { if(g != UNTYPED) typed = true; genotypes.push_back((unsigned char) g); }
    bool ismale() const { return gender == MALE; }
    bool isfemale() const { return gender == FEMALE; }
    bool isaffected() const { return affection_status == AFFECTED; }
    bool istyped() const { return typed; }
    bool mother_unknown() const { return mother == "0"; }
    bool father_unknown() const { return father == "0"; }
    bool isfounder_str() const { return mother_unknown() && father_unknown(); }
    bool isfounder() const { return maternal_id == UNKNOWN_PARENT && paternal_id == UNKNOWN_PARENT; }
    bool isleaf() const { return children.empty(); }
    bool is_parent(unsigned int i) const { return i == maternal_id || i == paternal_id; }
    bool operator<(const Person& p) const { return isfounder_str() && !p.isfounder_str(); }   // person.h:171-173
    double get_disease_prob(enum phased_trait pt) const { return disease_prob[pt]; }
    void make_unknown_affection(const DiseaseModel& dm) { affection_status = UNKNOWN_AFFECTION; init_probs(dm); }
    void clear_genotypes() { genotypes.clear(); }                                   // person.h:130-132
    // "experimental for the ELOD code" (person.h:202-208): the genotype prior at `locus` becomes the disease prior
    void copy_disease_probs(int locus) { for(int i = 0; i < 4; ++i) genotypes_prob[4 * locus + i] = disease_prob[i]; }
    void populate_trait_prob_cache(const GeneticMap& map, bool sex_linked);
    double get_trait_probability(unsigned int locus, enum phased_trait pt) const { return genotypes_prob[4 * locus + pt]; }
    bool safe_to_ignore_meiosis(const Pedigree& ped, enum parentage p, bool sex_linked) const;
};

class Pedigree {
    std::string id;
    bool sex_linked;
    std::vector<Person> members;
    unsigned int number_of_founders, number_of_leaves;
    bool mendelian_errors() const;
    int count_components() const;
 public:
    Pedigree(const std::string& id, bool sex_linked) : id(id), sex_linked(sex_linked), number_of_founders(0), number_of_leaves(0) {}
    std::string get_id() const { return id; }
    bool is_sexlinked() const { return sex_linked; }
    unsigned int num_members() const { return (unsigned int) members.size(); }
    unsigned int num_markers() const { return members[0].num_markers(); }
    unsigned int num_founders() const { return number_of_founders; }
    unsigned int num_leaves() const { return number_of_leaves; }
    Person* get_by_index(int i) { return &members[i]; }
    const Person* get_by_index(int i) const { return &members[i]; }
    Person* get_by_name(const std::string& id);
    bool add(const Person& p);
    bool exists(const std::string& id) { return get_by_name(id) != 0; }
    bool sanity_check();
};

// ---- parsers (parser.h, map_parser.cc, linkage_parser.cc, pedigree_parser.cc) ---------------
bool parse_map_file(const std::string& filename, GeneticMap& map);
bool parse_linkage_file(const std::string& filename, GeneticMap& map, DiseaseModel& dm);
bool parse_pedigree_file(const std::string& filename, std::vector<Pedigree>& pedigrees, const DiseaseModel& dm,
                         const GeneticMap& map, bool ignore_genotypes = false);     // PedigreeParser::set_ignore_genotypes
// Program::read_and_check_input (program.cc:18-63): map -> dat -> (force X) -> map sanity -> ped
bool read_and_check_input(const std::string& pedfile, const std::string& mapfile, const std::string& datfile,
                          bool force_sex_linked, GeneticMap& map, DiseaseModel& dm, std::vector<Pedigree>& pedigrees);

// ---- deterministic host RNG (replaces random.cc's GSL streams on the host side) --------------
class HostRng {
    uint64_t seed, counter;
 public:
    explicit HostRng(uint64_t seed) : seed(seed), counter(0) {}
    uint32_t next_u32();
    double uniform();
    int uniform_int(int n);
    template<typename T> void shuffle(std::vector<T>& v) {
        for(int i = (int) v.size() - 1; i > 0; --i) { int j = uniform_int(i + 1); T t = v[i]; v[i] = v[j]; v[j] = t; }
    }
};

// ---- genotype elimination (elimination.h/.cc) ------------------------------------------------
class DescentGraph;

class GenotypeElimination {
    Pedigree* ped;
    std::vector<int> possible;             // [M][N] masks
    bool init_processing;
    bool sex_linked;
    void initial_elimination();
    bool elimination_pass(int* ds);
    int child_homoz(int* ds, int mother, int father, int child, int homoz, bool ismale);
    int parent_homoz(int* ds, int parent, int other_parent, int child, int homoz, enum parentage p);
 public:
    GenotypeElimination(Pedigree* p, bool sex_linked) : ped(p), init_processing(false), sex_linked(sex_linked) {}
    bool elimination();
    bool random_descentgraph(DescentGraph& d, HostRng& rng);
    bool is_legal(int id, int locus, int value) const;
    int mask(int locus, int id) const { return possible[(size_t) locus * ped->num_members() + id]; }
};

// ---- peel planning (peeling.h, peel_sequence_generator.h/.cc) ---------------------------------
class PeelOperation {
    enum peeloperation type;
    unsigned int peelnode;
    bool used;
    std::vector<unsigned int> cutset, children, previous;
 public:
    explicit PeelOperation(unsigned int peelnode) : type(NULL_PEEL), peelnode(peelnode), used(false) {}
    unsigned int get_peelnode() const { return peelnode; }
    void set_used() { used = true; }
    bool is_used() const { return used; }
    bool in_cutset(unsigned int node) const;
    void add_cutnode(unsigned int c) { if(!in_cutset(c)) cutset.push_back(c); }
    void remove_cutnode(unsigned int c);
    unsigned int get_cutnode(unsigned int i) const { return cutset[i]; }
    unsigned int get_cutset_size() const { return (unsigned int) cutset.size(); }
    const std::vector<unsigned int>& get_cutset() const { return cutset; }
    bool contains_cutnodes(const std::vector<unsigned int>& nodes) const;
    unsigned int get_cost() const { return 1u << (2 * cutset.size()); }
    void set_type(enum peeloperation po, const Pedigree& ped);
    enum peeloperation get_type() const { return type; }
    const std::vector<unsigned int>& get_children() const { return children; }
    void add_prevfunction(unsigned int i) { previous.push_back(i); }
    const std::vector<unsigned int>& get_prevfunctions() const { return previous; }
    std::string translated_debug_string(const Pedigree& ped) const;
};

class PeelSequenceGenerator {
    Pedigree* ped;
    GeneticMap* map;
    bool verbose;
    std::vector<PeelOperation> peelorder;       // pedigree order before finalise, peel order after
    std::vector<PeelOperation> graph;           // the initial neighbour graph (build_simple_graph)
    std::vector<unsigned char> peeled;
    GenotypeElimination ge;
    HostRng rng;

    void build_simple_graph();
    void set_type(PeelOperation& p);
    void find_prev_functions(PeelOperation& op);
    static void eliminate_node(std::vector<PeelOperation>& tmp, unsigned int node);
    bool greedy_search(std::vector<unsigned int>& current);
    void random_downhill_search(std::vector<unsigned int>& current, unsigned int iterations);
 public:
    PeelSequenceGenerator(Pedigree* p, GeneticMap* m, bool sex_linked, bool verbose, uint64_t seed = 20261017);
    std::vector<PeelOperation>& get_peel_order() { return peelorder; }
    unsigned int get_peeling_cost() const;
    void build_peel_sequence(unsigned int iterations);
    // import an elimination order (e.g. one found by an earlier run); false if not legitimate
    bool set_peel_sequence(const std::vector<unsigned int>& seq);
    void finalise_peel_order(const std::vector<unsigned int>& seq);
    unsigned int get_cost(const std::vector<unsigned int>& seq) const;           // sum of cutset sizes
    unsigned int get_proper_cost(const std::vector<unsigned int>& seq) const;    // sum of 4^cutset
    bool is_legit(const std::vector<unsigned int>& seq) const;
    GenotypeElimination& get_elimination() { return ge; }
    std::string debug_string() const;
};

// ---- descent graph (descent_graph.h/.cc) -------------------------------------------------------
class DescentGraph {
    std::vector<int> data;
    Pedigree* ped;
    GeneticMap* map;
    double marker_transmission;
    int graph_size;
    bool sex_linked;
    int offset(unsigned person_id, unsigned locus, enum parentage p) const { return graph_size * (int) locus + 2 * (int) person_id + (int) p; }
 public:
    DescentGraph(Pedigree* ped, GeneticMap* map, bool sex_linked);
    int get(unsigned person_id, unsigned locus, enum parentage p) const { return data[offset(person_id, locus, p)]; }
    void set(unsigned person_id, unsigned locus, enum parentage p, int value) { data[offset(person_id, locus, p)] = value; }
    bool random_descentgraph(HostRng& rng);
    double get_marker_transmission() const { return marker_transmission; }
    double get_recombination_prob(unsigned int locus) const;
    int* get_internal_ptr() { return data.data(); }
    const int* get_internal_ptr() const { return data.data(); }
    size_t get_internal_size() const { return sizeof(int) * data.size(); }
};

// ---- LOD accumulator (lod_score.h) ----------------------------------------------------------------
class LODscores {
    GeneticMap* map;
    unsigned int num_scores_per_marker, num_scores, count;
    double trait_prob;
    std::vector<double> scores;
    std::vector<unsigned char> initialised;
 public:
    explicit LODscores(GeneticMap* map);
    void set_trait_prob(double prob) { trait_prob = prob; }
    double get_trait_prob() const { return trait_prob; }
    unsigned int num_lodscores() const { return num_scores; }
    unsigned int get_lodscores_per_marker() const { return num_scores_per_marker; }
    void add(unsigned int locus, unsigned int offset, double prob);
    double get_raw(unsigned int index) const { return scores[index]; }
    double get(unsigned int locus, unsigned int offset) const;
    double get_genetic_position(unsigned int locus, unsigned int offset) const { return map->get_genetic_position(locus, offset); }
    unsigned int get_count() const { return count; }
    void merge_results(LODscores* tmp);
    void set_count(unsigned int c) { count = c; }
    void set(unsigned int index, double prob) { scores[index] = prob; initialised[index] = 1; }
};

const double LOG_ILLEGAL = -1.7976931348623157e308;           // logarithms.h:10-11 (-DBL_MAX)
const double LOG_ZERO = LOG_ILLEGAL;
double log_sum(double a, double b);                           // logarithms.cc:14-23

// linkage_writer.cc:14-92
bool write_linkage_results(GeneticMap* map, const std::string& filename, std::vector<LODscores*>& all_scores, bool verbose);

// ---- flat problem for the C ABI ---------------------------------------------------------------------
struct FlatProblem {
    std::vector<int32_t> mother, father, sex, typed, prior_as_founder;
    std::vector<uint8_t> genotypes, elimination;
    std::vector<double> disease_prob, marker_prob, marker_xprob, theta, partial_theta, minor_freq, person_prior;
    std::vector<slk_peel_op> ops;
    slk_problem desc;
};
// fills `out` (and out.desc, pointing into out's vectors) from the host objects
// real_founder_priors: genotype priors computed with the resolved founder flags (the ELOD setup calls
// populate_trait_prob_cache after the pedigree is built, elod.h:76-83) instead of the parser's
// "everyone is a founder"; disease_prior_locus: Person::copy_disease_probs at that marker (-1: none)
void flatten_problem(Pedigree& ped, GeneticMap& map, PeelSequenceGenerator& psg, bool sex_linked, FlatProblem& out,
                     bool real_founder_priors = false, int disease_prior_locus = -1);

// ---- GPU back end (gpu_lodscores.h, gpu_markov_chain.h) -------------------------------------------------
class GPULodscores {
    Pedigree* ped;
    GeneticMap* map;
    PeelSequenceGenerator* psg;
    struct mcmc_options options;
    double trait_likelihood;
    FlatProblem flat;
    slk_plan* plan;
    slk_chain* chain;
    bool owns_plan;
 public:
    GPULodscores(Pedigree* ped, GeneticMap* map, PeelSequenceGenerator* psg, struct mcmc_options options, double trait_prob);
    ~GPULodscores();
    void calculate(DescentGraph& dg);           // gpu_lodscores.cc:598-607
    void block_until_finished();                // :609-619
    void get_results(LODscores* lod);           // :621-637
    slk_chain* get_chain() { return chain; }
    slk_plan* get_plan() { return plan; }
};

class GPUMarkovChain {
    Pedigree* ped;
    GeneticMap* map;
    PeelSequenceGenerator* psg;
    struct mcmc_options options;
    GeneticMap heated;          // the chain's own copy of the map (MarkovChain::map is a value member, heated in _init)
    double temperature;
    FlatProblem flat;
    slk_plan* plan;
    slk_chain* chain;
    int seq_num;
    double trait_prob;
    bool scoring_started;
    FILE* coda;
 public:
    // temperature as in MarkovChain(ped, map, psg, options, temp) (markov_chain.h:44-70): 1.0 = cold
    GPUMarkovChain(Pedigree* ped, GeneticMap* map, PeelSequenceGenerator* psg, struct mcmc_options options,
                   int sequence_num = 0, double temperature = 1.0);
    ~GPUMarkovChain();
    LODscores* run(DescentGraph& dg);           // gpu_markov_chain.cc:991 / markov_chain.cc:314
    // run() in three parts (every call only enqueues work on the chain's stream), so that several
    // replicate chains can be advanced in turn and overlap on the device: see run_replicates
    void begin(DescentGraph& dg);
    void iterate(int i);
    LODscores* finish(DescentGraph& dg);
    // MarkovChain::step (markov_chain.cc:107-207) on the chain's device-resident graph: step_size
    // iterations numbered from start_iteration; only the cold chain scores
    void step(int start_iteration, int step_size);
    void upload(DescentGraph& dg);
    void download(DescentGraph& dg);
    double get_likelihood();                    // MarkovChain::get_likelihood of the device-resident graph
    LODscores* get_result();                    // markov_chain.h:113-119
    double get_temperature() const { return temperature; }
    // SequentialImputation::parallel_run (sequential_imputation.cc:47-115) on the device: best of
    // `iterations` runs of LocusSampler::start_from, or locus_by_locus when iterations == 0
    double sequential_imputation(DescentGraph& dg, int iterations);
    double calc_trait_prob();                   // Peeler::calc_trait_prob on the device
    double get_likelihood(DescentGraph& dg);    // DescentGraph::get_likelihood on the device
    slk_chain* get_chain() { return chain; }
};

// the -R replicate loop of LinkageProgram::run_pedigree (linkage_program.cc:96-108), `in_flight` chains at a time
// on options.device; the caller owns the returned (merged) table
LODscores* run_replicates(Pedigree* ped, GeneticMap* map, PeelSequenceGenerator* psg, struct mcmc_options options, int in_flight);

// ---- ELOD (elod.h/.cc) ---------------------------------------------------------------------------------
// Expected LOD of a pedigree structure by simulation: a fake map marker - trait - marker; per replicate
// LocusSampler::start_from on the three loci (no genotypes; the trait locus's prior is each person's disease
// probability), the two marker rows scored by Peeler::process on the two-marker map.  All replicates of a
// pedigree run in one batch on the device (slk_elod_run).
class Elod {
    DiseaseModel dm;
    std::vector<Pedigree> pedigrees;
    GeneticMap map1, map2;
    struct mcmc_options options;
    std::vector<double> elods;
 public:
    Elod(const char* pedfile, struct mcmc_options opt);
    double run();                               // elod.cc:19-85: prints the table, returns the total
    const std::vector<double>& per_pedigree() const { return elods; }
};

// ---- Metropolis-coupled MCMC (mc3.h/.cc) ----------------------------------------------------------------
// The reference's Mc3 is compiled but unreachable (linkage_program.cc:169-170 is commented out, the
// -M/-z/-y/-t flags too, main.cc:70-75,382-428) and its driver numbers iterations i * spurts instead
// of i * exchange_period (mc3.cc:117).  This class specifies the intended behaviour: same ladder
// (mc3.cc:36-42), same heating (genetic_map.cc:95-117), same swap rule (mc3.cc:138-162), iterations
// numbered consecutively; every chain of the ladder lives on one device and a swap exchanges two
// device pointers.
double mc3_temperature(int chain_index, const struct mcmc_options& options);      // mc3.cc:31-42
class Mc3 {
    Pedigree* ped;
    GeneticMap* map;
    PeelSequenceGenerator* psg;
    struct mcmc_options options;
    std::vector<GPUMarkovChain*> chains;
    std::vector<int> swap_success, swap_failure;
    int seq_num;
    int period, spurts_done;
    HostRng rng;
 public:
    Mc3(Pedigree* ped, GeneticMap* map, PeelSequenceGenerator* psg, struct mcmc_options options, int sequence_num = 0);
    ~Mc3();
    LODscores* run();                           // mc3.cc:81-200; the caller deletes the result
    // run() = start(); total_spurts() x { enqueue_spurt(); exchange(); }; finish()
    void start();                               // start states, one sequential-imputation state per chain
    int total_spurts() const;
    int exchange_period() const { return period; }
    void enqueue_spurt();                       // `period` iterations of every chain (asynchronous)
    void exchange();                            // the Metropolis swap test of one adjacent pair (waits for the device)
    LODscores* finish();                        // the cold chain's table
    LODscores* cold_result() { return chains[0]->get_result(); }
    const std::vector<int>& get_swap_success() const { return swap_success; }
    const std::vector<int>& get_swap_failure() const { return swap_failure; }
};

// One device's share of a `-R` job (job.cc): the replicates placed on this device, plain chains or MC3 ladders,
// all resident at once and advanced in turn; results() merges their tables as LODscores::merge_results does.
class ReplicateJob {
    Pedigree* ped;
    GeneticMap* map;
    PeelSequenceGenerator* psg;
    struct mcmc_options options;
    std::vector<int> ids;
    bool ladder;
    std::vector<GPUMarkovChain*> chains;
    std::vector<DescentGraph*> dgs;
    std::vector<Mc3*> ladders;
    int next_iteration;
 public:
    ReplicateJob(Pedigree* ped, GeneticMap* map, PeelSequenceGenerator* psg, struct mcmc_options options,
                 const std::vector<int>& replicate_ids);
    ~ReplicateJob();
    int total_iterations() const;
    int done_iterations() const { return next_iteration; }
    int advance(int n);
    LODscores* results(std::vector<int>* swap_success, std::vector<int>* swap_failure);
};

}  // namespace swiftlink

#endif
