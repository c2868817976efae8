// peel.cc -- genotype elimination, peel-sequence generation, descent graph, LOD accumulator,
// output writer and the flattening of all of it into slk_problem.  Restates elimination.cc,
// peeling.h, peel_sequence_generator.cc, descent_graph.cc, lod_score.h, logarithms.cc and
// linkage_writer.cc of the reference.  For a given elimination order the derived peel
// operations (type, cutset order, previous functions, children) are bit-identical to the
// reference's; the search for an order uses this file's own counter-based RNG, because the
// reference's depends on GSL/libc streams that are not reproducible.
#include "swiftlink_host.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "../slk_philox.cuh"

namespace swiftlink {

// ---- host RNG ---------------------------------------------------------------------------

uint32_t HostRng::next_u32() {
    uint32_t o[4];
    slk_philox4x32_10((uint32_t) counter, (uint32_t)(counter >> 32), 0x484f5354u, 0u,
                      (uint32_t) seed, (uint32_t)(seed >> 32), o);
    ++counter;
    return o[0];
}

double HostRng::uniform() { return next_u32() / 4294967296.0; }

int HostRng::uniform_int(int n) {
    // rejection sampling, unbiased
    const uint32_t scale = 0xffffffffu / (uint32_t) n;
    uint32_t k;
    do { k = next_u32() / scale; } while(k >= (uint32_t) n);
    return (int) k;
}

// ---- genotype elimination (Lange-Goradia style, elimination.cc:41-125,127-240,360-395) -------

void GenotypeElimination::initial_elimination() {
    const unsigned N = ped->num_members(), M = ped->num_markers();
    possible.assign((size_t) M * N, GENO_AA | GENO_AB | GENO_BA | GENO_BB);
    for(unsigned l = 0; l < M; ++l) {
        for(unsigned j = 0; j < N; ++j) {
            const Person* p = ped->get_by_index(j);
            int& g = possible[(size_t) l * N + j];
            const enum unphased_genotype u = p->get_genotype(l);
            if(sex_linked && p->ismale() && u == UNTYPED) { g &= (GENO_AA | GENO_BB); continue; }
            switch(u) {
                case HOMOZ_A: g &= GENO_AA; break;
                case HETERO:  g &= (GENO_AB | GENO_BA); break;
                case HOMOZ_B: g &= GENO_BB; break;
                default: break;
            }
        }
    }
}

// elimination.cc:162-190
int GenotypeElimination::child_homoz(int* ds, int mother, int father, int child, int homoz, bool ismale) {
    const int other = (homoz == GENO_AA) ? GENO_BB : GENO_AA;
    int changes = 0;
    if(ds[child] == homoz) {
        if(ds[mother] & other) { ds[mother] ^= other; ++changes; }
        if(!(sex_linked && ismale)) {
            if(ds[father] & other) { ds[father] ^= other; ++changes; }
        }
    }
    return changes;
}

// elimination.cc:192-240
int GenotypeElimination::parent_homoz(int* ds, int parent, int other_parent, int child, int homoz, enum parentage p) {
    int changes = 0;
    if(ds[parent] == homoz) {
        int hetero, other_hetero;
        if((p == MATERNAL && homoz == GENO_AA) || (p == PATERNAL && homoz == GENO_BB)) { hetero = GENO_AB; other_hetero = GENO_BA; }
        else { hetero = GENO_BA; other_hetero = GENO_AB; }
        const int other_homoz = (homoz == GENO_AA) ? GENO_BB : GENO_AA;
        if(ds[child] & other_homoz)  { ds[child] ^= other_homoz; ++changes; }
        if(ds[child] & other_hetero) { ds[child] ^= other_hetero; ++changes; }
        if((ds[child] & hetero) && !(ds[child] & homoz)) {
            if(ds[other_parent] & homoz) { ds[other_parent] ^= homoz; ++changes; }
        }
    }
    return changes;
}

// elimination.cc:87-125; ds = masks of one locus
bool GenotypeElimination::elimination_pass(int* ds) {
    const unsigned N = ped->num_members();
    while(true) {
        int changes = 0;
        for(unsigned i = 0; i < N; ++i) {
            const Person* p = ped->get_by_index(i);
            if(p->isfounder()) continue;
            const int pat = p->get_paternalid(), mat = p->get_maternalid();
            changes += parent_homoz(ds, mat, pat, i, GENO_AA, MATERNAL);
            changes += parent_homoz(ds, mat, pat, i, GENO_BB, MATERNAL);
            if(!(sex_linked && p->ismale())) {
                changes += parent_homoz(ds, pat, mat, i, GENO_AA, PATERNAL);
                changes += parent_homoz(ds, pat, mat, i, GENO_BB, PATERNAL);
            }
            changes += child_homoz(ds, mat, pat, i, GENO_AA, p->ismale());
            changes += child_homoz(ds, mat, pat, i, GENO_BB, p->ismale());
        }
        for(unsigned i = 0; i < N; ++i) if(ds[i] == 0) return false;
        if(changes == 0) break;
    }
    return true;
}

bool GenotypeElimination::elimination() {
    if(!init_processing) {
        initial_elimination();
        const unsigned N = ped->num_members();
        for(unsigned l = 0; l < ped->num_markers(); ++l) {
            if(!elimination_pass(&possible[(size_t) l * N])) return false;
        }
        init_processing = true;
    }
    return true;
}

// genotype.cc:91-102 + elimination.cc:393-395
bool GenotypeElimination::is_legal(int id, int locus, int value) const {
    int m = 0;
    switch(value) {
        case TRAIT_UU: m = GENO_AA; break;
        case TRAIT_AU: m = GENO_BA; break;
        case TRAIT_UA: m = GENO_AB; break;
        case TRAIT_AA: m = GENO_BB; break;
        default:
            fprintf(stderr, "error in genotype_from_trait, %d\n", value);
            abort();
    }
    return (possible[(size_t) locus * ped->num_members() + id] & m) != 0;
}

// elimination.cc:242-358: pick random single genotypes consistent with elimination, locus by
// locus, then read meiosis indicators off them
bool GenotypeElimination::random_descentgraph(DescentGraph& d, HostRng& rng) {
    if(!elimination()) return false;
    const unsigned N = ped->num_members(), M = ped->num_markers();
    std::vector<int> ds(N), queue(N);
    for(unsigned l = 0; l < M; ++l) {
        std::copy(possible.begin() + (size_t) l * N, possible.begin() + (size_t)(l + 1) * N, ds.begin());
        while(true) {
            if(!elimination_pass(ds.data())) {
                std::copy(possible.begin() + (size_t) l * N, possible.begin() + (size_t)(l + 1) * N, ds.begin());
            }
            bool complete = true;
            for(unsigned i = 0; i < N; ++i) if((ds[i] & (ds[i] - 1)) != 0) { complete = false; break; }
            if(complete) break;
            int q = 0;
            for(unsigned i = 0; i < N; ++i) if((ds[i] & (ds[i] - 1)) != 0) queue[q++] = i;
            const int person = queue[rng.uniform_int(q)];
            q = 0;
            for(int b = 0; b < 4; ++b) if(ds[person] & (1 << b)) queue[q++] = 1 << b;
            ds[person] = queue[rng.uniform_int(q)];
        }
        for(unsigned j = 0; j < N; ++j) {
            const Person* p = ped->get_by_index(j);
            if(p->isfounder()) { d.set(j, l, MATERNAL, 0); d.set(j, l, PATERNAL, 0); continue; }
            const int c = ds[j];
            for(int par = 0; par < 2; ++par) {
                const int pg = ds[par == 0 ? p->get_maternalid() : p->get_paternalid()];
                const int hetero = (par == 0) ? GENO_AB : GENO_BA;      // elimination.cc:262-290
                int state;
                if(c == GENO_AA || c == hetero) state = (pg == GENO_AB) ? 0 : (pg == GENO_BA) ? 1 : rng.uniform_int(2);
                else                            state = (pg == GENO_BA) ? 0 : (pg == GENO_AB) ? 1 : rng.uniform_int(2);
                d.set(j, l, (enum parentage) par, state);
            }
        }
    }
    return true;
}

// ---- peel operations (peeling.h) ----------------------------------------------------------------

bool PeelOperation::in_cutset(unsigned int node) const {
    return std::find(cutset.begin(), cutset.end(), node) != cutset.end();
}

void PeelOperation::remove_cutnode(unsigned int c) {
    std::vector<unsigned int>::iterator it = std::find(cutset.begin(), cutset.end(), c);
    if(it != cutset.end()) cutset.erase(it);
}

bool PeelOperation::contains_cutnodes(const std::vector<unsigned int>& nodes) const {
    for(size_t i = 0; i < cutset.size(); ++i)
        if(std::find(nodes.begin(), nodes.end(), cutset[i]) == nodes.end()) return false;
    return true;
}

// peeling.h:153-171
void PeelOperation::set_type(enum peeloperation po, const Pedigree& ped) {
    type = po;
    if(type == PARENT_PEEL) {
        const Person* p = ped.get_by_index(peelnode);
        for(size_t i = 0; i < cutset.size(); ++i) if(p->is_offspring(cutset[i])) children.push_back(cutset[i]);
    }
}

std::string PeelOperation::translated_debug_string(const Pedigree& ped) const {
    static const char* names[] = {"null", "child", "parent", "partner", "last"};
    std::ostringstream ss;
    ss << names[type] << "\t" << "peelnode = " << peelnode << "\t" << "id = " << ped.get_by_index(peelnode)->get_id() << "\t" << "cutset = (";
    for(size_t i = 0; i < cutset.size(); ++i) ss << ped.get_by_index(cutset[i])->get_id() << (i + 1 != cutset.size() ? "," : "");
    ss << ")\tprev = (";
    for(size_t i = 0; i < previous.size(); ++i) ss << previous[i] << (i + 1 != previous.size() ? "," : "");
    ss << ")  children = (";
    for(size_t i = 0; i < children.size(); ++i) ss << ped.get_by_index(children[i])->get_id() << (i + 1 != children.size() ? "," : "");
    ss << ") ";
    return ss.str();
}

// ---- peel sequence generator -----------------------------------------------------------------------

PeelSequenceGenerator::PeelSequenceGenerator(Pedigree* p, GeneticMap* m, bool sex_linked, bool verbose, uint64_t seed) :
    ped(p), map(m), verbose(verbose), peeled(p->num_members(), 0), ge(p, sex_linked), rng(seed) {
    ge.elimination();
    build_simple_graph();
}

// peel_sequence_generator.cc:245-271: neighbours = parents, children, mates, in that order
void PeelSequenceGenerator::build_simple_graph() {
    graph.clear();
    for(unsigned int i = 0; i < ped->num_members(); ++i) {
        PeelOperation po(i);
        const Person* p = ped->get_by_index(i);
        if(!p->isfounder()) {
            po.add_cutnode(p->get_maternalid());
            po.add_cutnode(p->get_paternalid());
        }
        for(unsigned int j = 0; j < p->num_children(); ++j) po.add_cutnode(p->get_child(j));
        for(unsigned int j = 0; j < p->num_mates(); ++j) po.add_cutnode(p->get_mate(j));
        graph.push_back(po);
    }
    peelorder = graph;
}

// peel_sequence_generator.cc:411-423
void PeelSequenceGenerator::eliminate_node(std::vector<PeelOperation>& tmp, unsigned int node) {
    const std::vector<unsigned int> cutset = tmp[node].get_cutset();
    for(size_t i = 0; i < cutset.size(); ++i) {
        tmp[cutset[i]].remove_cutnode(node);
        for(size_t j = 0; j < cutset.size(); ++j) {
            if(i != j) {
                tmp[cutset[i]].add_cutnode(cutset[j]);
                tmp[cutset[j]].add_cutnode(cutset[i]);
            }
        }
    }
}

// peel_sequence_generator.cc:188-223
void PeelSequenceGenerator::set_type(PeelOperation& p) {
    const Person* q = ped->get_by_index(p.get_peelnode());
    enum peeloperation t;
    bool final_node = true;
    for(unsigned int i = 0; i < peeled.size(); ++i) if(i != p.get_peelnode() && !peeled[i]) { final_node = false; break; }
    unsigned unpeeled_mates = 0, unpeeled_children = 0;
    for(unsigned int i = 0; i < q->num_mates(); ++i) if(!peeled[q->get_mate(i)]) ++unpeeled_mates;
    for(unsigned int i = 0; i < q->num_children(); ++i) if(!peeled[q->get_child(i)]) ++unpeeled_children;

    if(final_node) t = LAST_PEEL;
    else if(!q->isfounder() && !(peeled[q->get_maternalid()] || peeled[q->get_paternalid()])) t = CHILD_PEEL;
    else if(!q->isleaf() && unpeeled_mates != 0 && unpeeled_children != 0) t = PARENT_PEEL;
    else t = PARTNER_PEEL;
    p.set_type(t, *ped);
}

// peel_sequence_generator.cc:56-82
void PeelSequenceGenerator::find_prev_functions(PeelOperation& op) {
    std::vector<unsigned int> nodes(op.get_cutset());
    nodes.push_back(op.get_peelnode());
    while(true) {
        int found = -1;
        for(int i = 0; i < int(peelorder.size()); ++i) {
            if(peelorder[i].is_used()) continue;
            if(peelorder[i].contains_cutnodes(nodes)) { peelorder[i].set_used(); found = i; break; }
        }
        if(found == -1) break;
        op.add_prevfunction(found);
    }
}

// peel_sequence_generator.cc:225-243 (the per-locus index tables of bruteforce_assignments are
// not built: validity is derived on the device from the elimination masks)
void PeelSequenceGenerator::finalise_peel_order(const std::vector<unsigned int>& seq) {
    std::vector<PeelOperation> tmp(graph);
    peelorder.clear();
    std::fill(peeled.begin(), peeled.end(), 0);
    for(size_t i = 0; i < seq.size(); ++i) {
        PeelOperation p = tmp[seq[i]];
        set_type(p);
        find_prev_functions(p);
        eliminate_node(tmp, seq[i]);
        peelorder.push_back(p);
        peeled[seq[i]] = 1;
    }
}

// Cost functions and legitimacy work on adjacency bitsets: the order inside a cutset does not
// matter for them, and one evaluation is O(N * c * N/64) instead of the reference's vector
// surgery (340 s per 10^6 search iterations at N = 200, probe).
namespace {
struct BitGraph {
    int n, words;
    std::vector<uint64_t> adj;
    BitGraph(const std::vector<PeelOperation>& g) : n((int) g.size()), words(((int) g.size() + 63) / 64), adj((size_t) n * words, 0) {
        for(int i = 0; i < n; ++i)
            for(unsigned k = 0; k < g[i].get_cutset_size(); ++k) set(i, g[i].get_cutnode(k));
    }
    void set(int i, int j) { adj[(size_t) i * words + (j >> 6)] |= 1ull << (j & 63); }
    void clear(int i, int j) { adj[(size_t) i * words + (j >> 6)] &= ~(1ull << (j & 63)); }
    int degree(int i) const { int d = 0; for(int w = 0; w < words; ++w) d += __builtin_popcountll(adj[(size_t) i * words + w]); return d; }
    void eliminate(int v) {
        const uint64_t* s = &adj[(size_t) v * words];
        for(int w = 0; w < words; ++w) {
            uint64_t bits = s[w];
            while(bits) {
                int u = (w << 6) + __builtin_ctzll(bits);
                bits &= bits - 1;
                uint64_t* a = &adj[(size_t) u * words];
                for(int x = 0; x < words; ++x) a[x] |= s[x];
                clear(u, u);
                clear(u, v);
            }
        }
    }
};
}  // namespace

unsigned int PeelSequenceGenerator::get_cost(const std::vector<unsigned int>& seq) const {
    BitGraph g(graph);
    unsigned int cost = 0;
    for(size_t i = 0; i < seq.size(); ++i) { cost += g.degree(seq[i]); g.eliminate(seq[i]); }
    return cost;
}

unsigned int PeelSequenceGenerator::get_proper_cost(const std::vector<unsigned int>& seq) const {
    BitGraph g(graph);
    unsigned int cost = 0;
    for(size_t i = 0; i < seq.size(); ++i) { cost += 1u << (2 * g.degree(seq[i])); g.eliminate(seq[i]); }
    return cost;
}

// peel_sequence_generator.cc:453-483
bool PeelSequenceGenerator::is_legit(const std::vector<unsigned int>& seq) const {
    std::vector<unsigned char> done(ped->num_members(), 0);
    if(seq.size() != ped->num_members()) return false;
    for(size_t i = 0; i < seq.size(); ++i) {
        if(seq[i] >= ped->num_members() || done[seq[i]]) return false;
        const Person* q = ped->get_by_index(seq[i]);
        unsigned um = 0, uc = 0;
        for(unsigned int k = 0; k < q->num_mates(); ++k) if(!done[q->get_mate(k)]) ++um;
        for(unsigned int k = 0; k < q->num_children(); ++k) if(!done[q->get_child(k)]) ++uc;
        if(!q->isfounder() && !(done[q->get_maternalid()] || done[q->get_paternalid()]) &&
           !q->isleaf() && um != 0 && uc != 0) return false;
        done[seq[i]] = 1;
    }
    return true;
}

// peel_sequence_generator.cc:309-359: succeeds only when every elimination has <= 2 neighbours
bool PeelSequenceGenerator::greedy_search(std::vector<unsigned int>& current) {
    BitGraph g(graph);
    std::vector<unsigned char> done(graph.size(), 0);
    current.clear();
    while(current.size() != graph.size()) {
        int best = 1 << 30;
        std::vector<unsigned int> ties;
        for(unsigned int i = 0; i < graph.size(); ++i) {
            if(done[i]) continue;
            int d = g.degree(i);
            if(d < best) { best = d; ties.clear(); }
            if(d == best) ties.push_back(i);
        }
        if(best > 2) {
            fprintf(stderr, "Greedy algorithm to find peeling sequence failed, switching to randomised method...\n");
            return false;
        }
        unsigned int node = ties[rng.uniform_int((int) ties.size())];
        g.eliminate(node);
        done[node] = 1;
        current.push_back(node);
    }
    return true;
}

// peel_sequence_generator.cc:361-409
void PeelSequenceGenerator::random_downhill_search(std::vector<unsigned int>& current, unsigned int iterations) {
    unsigned int cost = get_cost(current);
    for(unsigned int i = 0; i < iterations; ++i) {
        int a, b;
        do { a = rng.uniform_int((int) current.size()); b = rng.uniform_int((int) current.size()); } while(a == b);
        std::swap(current[a], current[b]);
        unsigned int new_cost = get_cost(current);
        if(new_cost <= cost) { cost = new_cost; continue; }
        std::swap(current[a], current[b]);
    }
    if(verbose) fprintf(stderr, "Peel Sequence: cost = %d\n", get_proper_cost(current));
}

// peel_sequence_generator.cc:273-307
void PeelSequenceGenerator::build_peel_sequence(unsigned int iterations) {
    std::vector<unsigned int> current;
    if(greedy_search(current) && is_legit(current)) {
        finalise_peel_order(current);
        return;
    }
    while(true) {
        current.clear();
        for(unsigned i = 0; i < ped->num_members(); ++i) current.push_back(i);
        rng.shuffle(current);
        random_downhill_search(current, iterations);
        if(is_legit(current)) break;
    }
    finalise_peel_order(current);
}

bool PeelSequenceGenerator::set_peel_sequence(const std::vector<unsigned int>& seq) {
    if(!is_legit(seq)) return false;
    finalise_peel_order(seq);
    return true;
}

unsigned int PeelSequenceGenerator::get_peeling_cost() const {
    unsigned int cost = 0;
    for(size_t i = 0; i < peelorder.size(); ++i) cost += peelorder[i].get_cost();
    return cost;
}

std::string PeelSequenceGenerator::debug_string() const {
    std::ostringstream ss;
    for(size_t i = 0; i < peelorder.size(); ++i) ss << i << "\t" << peelorder[i].translated_debug_string(*ped) << "\n";
    return ss.str();
}

// ---- descent graph ------------------------------------------------------------------------------------

DescentGraph::DescentGraph(Pedigree* ped, GeneticMap* map, bool sex_linked) :
    data((size_t) 2 * ped->num_members() * map->num_markers(), 0), ped(ped), map(map),
    marker_transmission(log(0.5) * (2 * (ped->num_members() - ped->num_founders()))),      // descent_graph.cc:22
    graph_size(2 * (int) ped->num_members()), sex_linked(sex_linked) {
    if(sex_linked) marker_transmission = log(0.5) * (ped->num_members() - ped->num_founders());   // :35
}

bool DescentGraph::random_descentgraph(HostRng& rng) {
    GenotypeElimination ge(ped, sex_linked);
    return ge.random_descentgraph(*this, rng);
}

// descent_graph.cc:212-242
double DescentGraph::get_recombination_prob(unsigned int locus) const {
    double tmp = 0.0;
    const double theta = map->get_theta_log(locus), antitheta = map->get_inversetheta_log(locus);
    const unsigned num_alleles = sex_linked ? 1 : 2;
    for(unsigned i = 0; i < ped->num_members(); ++i) {
        if(ped->get_by_index(i)->isfounder()) continue;
        for(unsigned j = 0; j < num_alleles; ++j) {
            const enum parentage parent = (enum parentage) j;
            tmp += (get(i, locus, parent) != get(i, locus + 1, parent)) ? theta : antitheta;
        }
    }
    return tmp;
}

// ---- LOD accumulator ---------------------------------------------------------------------------------------

double log_sum(double a, double b) {
    if(a == -DBL_MAX) return b;
    if(b == -DBL_MAX) return a;
    return log(exp(b - a) + 1) + a;
}

LODscores::LODscores(GeneticMap* map) :
    map(map), num_scores_per_marker(map->get_lodscore_count()),
    num_scores(map->get_lodscore_count() * (map->num_markers() - 1)), count(0), trait_prob(0.0),
    scores(num_scores, 0.0), initialised(num_scores, 0) {}

// lod_score.h:74-80
void LODscores::add(unsigned int locus, unsigned int offset, double prob) {
    const unsigned int index = locus * num_scores_per_marker + offset;
    scores[index] = initialised[index] ? log_sum(prob, scores[index]) : prob;
    initialised[index] = 1;
    if(locus == 0 && offset == 0) ++count;
}

// lod_score.h:86-88
double LODscores::get(unsigned int locus, unsigned int offset) const {
    return (scores[locus * num_scores_per_marker + offset] - log((double) count) - trait_prob) / log(10.0);
}

// lod_score.h:98-105
void LODscores::merge_results(LODscores* tmp) {
    for(unsigned i = 0; i < num_scores; ++i) scores[i] = log_sum(scores[i], tmp->get_raw(i));
    count += tmp->get_count();
}

// linkage_writer.cc:52-84
bool write_linkage_results(GeneticMap* map, const std::string& filename, std::vector<LODscores*>& all_scores, bool verbose) {
    std::ofstream f(filename.c_str(), std::ios::out | std::ios::trunc);
    if(!f.is_open()) {
        fprintf(stderr, "error: could not open linkage output file \"%s\"\n", filename.c_str());
        return false;
    }
    f << "marker\tposition\tlod\n";
    for(unsigned int i = 0; i < map->num_markers() - 1; ++i) {
        f << map->get_name(i) << "\t" << 100.0 * map->get_genetic_position(i, 0) << "\n";
        for(unsigned int j = 0; j < map->get_lodscore_count(); ++j) {
            double total = all_scores[0]->get(i, j);
            for(size_t k = 1; k < all_scores.size(); ++k) total += all_scores[k]->get(i, j);
            std::ostringstream ss;
            ss << "-\t" << 100 * map->get_genetic_position(i, j + 1) << "\t" << total;
            if(all_scores.size() > 1) for(size_t k = 0; k < all_scores.size(); ++k) ss << "\t" << all_scores[k]->get(i, j);
            ss << "\n";
            f << ss.str();
            if(verbose) fprintf(stderr, "%s", ss.str().c_str());
        }
    }
    f << map->get_name(map->num_markers() - 1) << "\t" << 100.0 * map->get_genetic_position(map->num_markers() - 1, 0) << "\n";
    f.close();
    return true;
}

// ---- flat problem -----------------------------------------------------------------------------------------------

void flatten_problem(Pedigree& ped, GeneticMap& map, PeelSequenceGenerator& psg, bool sex_linked, FlatProblem& out,
                     bool real_founder_priors, int disease_prior_locus) {
    const int N = (int) ped.num_members(), M = (int) map.num_markers();
    out.mother.resize(N); out.father.resize(N); out.sex.resize(N); out.typed.resize(N); out.prior_as_founder.resize(N);
    out.genotypes.resize((size_t) N * M);
    out.disease_prob.resize((size_t) N * 4);
    for(int i = 0; i < N; ++i) {
        const Person* p = ped.get_by_index(i);
        out.mother[i] = p->isfounder() ? -1 : (int) p->get_maternalid();
        out.father[i] = p->isfounder() ? -1 : (int) p->get_paternalid();
        out.sex[i] = (int) p->get_sex();
        out.typed[i] = p->istyped() ? 1 : 0;
        out.prior_as_founder[i] = p->isfounder() ? 1 : 0;
        for(int l = 0; l < M; ++l) out.genotypes[(size_t) i * M + l] = (uint8_t) p->get_marker(l);
        for(int g = 0; g < 4; ++g) out.disease_prob[(size_t) i * 4 + g] = p->get_disease_prob((enum phased_trait) g);
    }
    out.marker_prob.resize((size_t) M * 4); out.marker_xprob.resize((size_t) M * 4);
    out.theta.resize(M - 1); out.partial_theta.resize(M - 1);
    out.minor_freq.resize(M);
    for(int l = 0; l < M; ++l) {
        out.minor_freq[l] = map.get_minor(l);
        for(int g = 0; g < 4; ++g) {
            out.marker_prob[(size_t) l * 4 + g] = map.get_prob(l, (enum phased_trait) g, false);
            out.marker_xprob[(size_t) l * 4 + g] = map.get_prob(l, (enum phased_trait) g, true);
        }
    }
    for(int l = 0; l < M - 1; ++l) { out.theta[l] = map.get_theta(l); out.partial_theta[l] = map.get_theta_partial_raw(l); }
    out.elimination.resize((size_t) M * N);
    GenotypeElimination& ge = psg.get_elimination();
    for(int l = 0; l < M; ++l) for(int i = 0; i < N; ++i) out.elimination[(size_t) l * N + i] = (uint8_t) ge.mask(l, i);

    std::vector<PeelOperation>& ops = psg.get_peel_order();
    out.ops.resize(ops.size());
    for(size_t i = 0; i < ops.size(); ++i) {
        slk_peel_op& o = out.ops[i];
        memset(&o, 0, sizeof(o));
        const PeelOperation& p = ops[i];
        if(p.get_cutset_size() > SLK_MAX_CUTSET || p.get_prevfunctions().size() > SLK_MAX_PREV || p.get_children().size() > SLK_MAX_CHILDREN) {
            fprintf(stderr, "error: peel operation %d exceeds the device plan limits (cutset %d, previous %d, children %d)\n",
                    (int) i, (int) p.get_cutset_size(), (int) p.get_prevfunctions().size(), (int) p.get_children().size());
            abort();
        }
        o.type = (int32_t) p.get_type();
        o.peelnode = (int32_t) p.get_peelnode();
        o.ncut = (int32_t) p.get_cutset_size();
        for(int k = 0; k < o.ncut; ++k) o.cutset[k] = (int32_t) p.get_cutnode(k);
        o.nprev = (int32_t) p.get_prevfunctions().size();
        for(int k = 0; k < o.nprev; ++k) o.prev[k] = (int32_t) p.get_prevfunctions()[k];
        o.nchild = (int32_t) p.get_children().size();
        for(int k = 0; k < o.nchild; ++k) o.children[k] = (int32_t) p.get_children()[k];
    }

    slk_problem& d = out.desc;
    memset(&d, 0, sizeof(d));
    d.n_members = N; d.n_founders = (int) ped.num_founders(); d.n_markers = M; d.n_lod = (int) map.get_lodscore_count();
    d.sex_linked = sex_linked ? 1 : 0;
    d.mother = out.mother.data(); d.father = out.father.data(); d.sex = out.sex.data(); d.typed = out.typed.data();
    d.prior_as_founder = real_founder_priors ? out.prior_as_founder.data() : 0;     // NULL: reference behaviour, see include/swiftlink_b200.h
    d.disease_prior_locus_plus1 = disease_prior_locus + 1;
    if(disease_prior_locus >= 0) {
        // what Person::copy_disease_probs stored at that marker -- NOT get_disease_prob(): make_unknown_affection may
        // have re-run init_probs since the copy (elod.h:83-87)
        out.person_prior.resize((size_t) N * 4);
        for(int i = 0; i < N; ++i)
            for(int g = 0; g < 4; ++g)
                out.person_prior[(size_t) i * 4 + g] = ped.get_by_index(i)->get_trait_probability(disease_prior_locus, (enum phased_trait) g);
        d.person_prior = out.person_prior.data();
    }
    d.genotypes = out.genotypes.data(); d.disease_prob = out.disease_prob.data();
    d.marker_prob = out.marker_prob.data(); d.marker_xprob = out.marker_xprob.data();
    d.theta = out.theta.data(); d.partial_theta = out.partial_theta.data();
    d.elimination = out.elimination.data();
    d.n_ops = (int) out.ops.size(); d.ops = out.ops.data();
    d.minor_freq = out.minor_freq.data();
}

}  // namespace swiftlink
