// model.cc -- DiseaseModel, Snp/GeneticMap, Person, Pedigree and the LINKAGE ped/map/dat
// readers.  Restates disease_model.cc, genetic_map.cc, person.cc, pedigree.cc,
// pedigree_parser.cc, map_parser.cc, linkage_parser.cc and parser.h of the reference so that
// every table the hot path consumes is bit-identical (tests/test_host.py).
//
// This file is the drop-in boundary's host mirror, not the hot path: class names, messages and the arithmetic of the
// tables have to be the reference's.  Three bodies follow their originals statement for statement because every
// statement is either an error message a user greps for or an operation whose order fixes a table bit for bit:
// GeneticMap::sanity_check (genetic_map.cc:14-64), GeneticMap::set_temperature (:95-117) and Pedigree::sanity_check
// (pedigree.cc); they are marked where they stand.
#include "swiftlink_host.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <queue>
#include <sstream>

namespace swiftlink {

mcmc_options::mcmc_options() :
    verbose(false), burnin(50000), iterations(50000), si_iterations(1000), scoring_period(10), mcmc_runs(1),
    coda_logging(false), coda_prefix("trace"), lodscores(5), peelopt_iterations(1000000), lsampler_prob(0.5),
    thread_count(4), use_gpu(true), random_filename(""), affected_only(false), sex_linked(false),
    seed(20261017), device(0), elod(false), elod_frequency(0.0001), elod_marker_separation(0.05), elod_replicates(1000000),
    mc3(false), mc3_number_of_chains(1), mc3_exchange_period(10),
    mc3_temperatures(), exchange_filename("") {                    // defaults.h:4-24
    elod_penetrance[0] = 0.0; elod_penetrance[1] = 0.0; elod_penetrance[2] = 1.0;
}

// ---- disease model -----------------------------------------------------------------------

// disease_model.cc:139-153
double DiseaseModel::get_penetrance_prob2(enum affection a, enum unphased_trait t, enum sex s) const {
    const bool xmale = sexlinked && s == MALE;
    if(xmale && t == TRAIT_HETERO) return 0.0;
    if(a == UNKNOWN_AFFECTION) return xmale ? 0.5 : 0.25;
    return (a == AFFECTED) ? penetrance[t] : 1.0 - penetrance[t];
}

// disease_model.cc:155-187
double DiseaseModel::get_apriori_prob2(enum affection a, enum unphased_trait t, enum sex s) const {
    const bool xmale = sexlinked && s == MALE;
    double tmp = 0.0;
    if(xmale && t == TRAIT_HETERO) return 0.0;
    switch(t) {
        case TRAIT_HOMO_A: tmp = xmale ? frequency : frequency * frequency; break;
        case TRAIT_HETERO: tmp = xmale ? 0.0 : frequency * (1.0 - frequency); break;
        case TRAIT_HOMO_U: tmp = xmale ? (1.0 - frequency) : (1.0 - frequency) * (1.0 - frequency); break;
    }
    switch(a) {
        case AFFECTED:          return tmp * penetrance[t];
        case UNAFFECTED:        return tmp * (1.0 - penetrance[t]);
        case UNKNOWN_AFFECTION: return (tmp * penetrance[t]) + (tmp * (1.0 - penetrance[t]));
    }
    abort();
}

// ---- genetic map ------------------------------------------------------------------------------

// genetic_map.h:68-91
void Snp::init_probs() {
    prob[TRAIT_UU] = major_freq * major_freq;
    prob[TRAIT_AU] = prob[TRAIT_UA] = minor_freq * major_freq;
    prob[TRAIT_AA] = minor_freq * minor_freq;
    double total = prob[TRAIT_UU] + prob[TRAIT_AU] + prob[TRAIT_UA] + prob[TRAIT_AA];
    for(int i = 0; i < 4; ++i) prob[i] /= total;

    x_male_prob[TRAIT_UU] = major_freq;
    x_male_prob[TRAIT_UA] = x_male_prob[TRAIT_AU] = 0.0;
    x_male_prob[TRAIT_AA] = minor_freq;
    total = x_male_prob[TRAIT_UU] + x_male_prob[TRAIT_UA] + x_male_prob[TRAIT_AU] + x_male_prob[TRAIT_AA];
    for(int i = 0; i < 4; ++i) x_male_prob[i] /= total;
}

double GeneticMap::haldane(double m) const { return 0.5 * (1.0 - exp(-2.0 * m)); }           // genetic_map.cc:86-88
double GeneticMap::inverse_haldane(double r) const { return -0.5 * log(1 - (2 * r)); }        // :90-92
double GeneticMap::get_theta_log(unsigned int i) const { return log(thetas[i]); }
double GeneticMap::get_inversetheta_log(unsigned int i) const { return log(inversethetas[i]); }

double GeneticMap::get_genetic_position(unsigned int index, unsigned int offset) const {      // :119-121
    return map[index].get_g_distance() + (offset ? inverse_haldane(partial_thetas[index] * offset) : 0);
}

// genetic_map.cc:14-64
bool GeneticMap::sanity_check() {
    if(map.size() != thetas.size() + 1) {
        bool dat_warning = false;
        fprintf(stderr, "Error: number of markers = %d, number of thetas = %d\n", int(map.size()), int(thetas.size()));
        for(int i = 0; i < int(map.size()); ++i) {
            if(!map[i].is_maf_set()) {
                fprintf(stderr, "Error: marker %d (%s) does not have a minor allele frequency\n", i + 1, map[i].get_name().c_str());
                dat_warning = true;
            }
        }
        if(dat_warning) fprintf(stderr, "Error: please check that DAT and MAP files are consistent...\n");
        return false;
    }
    for(unsigned i = 0; i < map.size(); ++i) map[i].init_probs();

    // the .dat recombination fractions are discarded and recomputed from the map positions
    fprintf(stderr, "WARNING: recalculating theta values using Haldane map function\n");
    thetas.clear();
    inversethetas.clear();
    for(unsigned i = 1; i < map.size(); ++i) {
        double tmp = haldane(map[i].get_g_distance() - map[i-1].get_g_distance());
        add_theta(tmp);
        if(tmp == 0.0) {
            fprintf(stderr, "Error: recombination fraction between %s and %s (markers %d and %d) is zero! (sampling will not work properly)\nExiting...\n",
                    map[i-1].get_name().c_str(), map[i].get_name().c_str(), i - 1, i);
            exit(EXIT_FAILURE);
        }
    }
    partial_thetas.clear();
    for(unsigned i = 0; i < thetas.size(); ++i) {
        partial_thetas.push_back(haldane(inverse_haldane(thetas[i]) / double(partial_theta_count + 1)));
    }
    return true;
}

// genetic_map.cc:95-117 (MC3 heating: 1.0 = cold, 0.0 = uniform)
void GeneticMap::set_temperature(double t) {
    if(temperature != 1.0) {
        fprintf(stderr, "error: temperature cannot be set twice in GeneticMap objects\n");
        abort();
    }
    temperature = t;
    for(unsigned i = 0; i < thetas.size(); ++i) {
        thetas[i] = (temperature * thetas[i]) + ((1 - temperature) * 0.5);
        inversethetas[i] = 1.0 - thetas[i];
    }
    for(unsigned i = 0; i < map.size(); ++i) {
        double minor_freq = map[i].minor();
        map[i].set_minor_freq((temperature * minor_freq) + ((1 - temperature) * 0.5));
    }
}

// ---- person ---------------------------------------------------------------------------------------

Person::Person(const std::string& name, const std::string& father_name, const std::string& mother_name,
               enum sex s, enum affection a, const DiseaseModel& dm) :
    id(name), mother(mother_name), father(father_name), gender(s), affection_status(a),
    internal_id(UNKNOWN_PARENT), maternal_id(UNKNOWN_PARENT), paternal_id(UNKNOWN_PARENT), typed(false) {
    init_probs(dm);
}

// person.cc:85-119
void Person::init_probs(const DiseaseModel& dm) {
    const bool f = isfounder_str();
    disease_prob[TRAIT_AA] = f ? dm.get_apriori_prob2(affection_status, TRAIT_HOMO_A, gender)
                               : dm.get_penetrance_prob2(affection_status, TRAIT_HOMO_A, gender);
    disease_prob[TRAIT_AU] = disease_prob[TRAIT_UA] =
                             f ? dm.get_apriori_prob2(affection_status, TRAIT_HETERO, gender)
                               : dm.get_penetrance_prob2(affection_status, TRAIT_HETERO, gender);
    disease_prob[TRAIT_UU] = f ? dm.get_apriori_prob2(affection_status, TRAIT_HOMO_U, gender)
                               : dm.get_penetrance_prob2(affection_status, TRAIT_HOMO_U, gender);
}

bool Person::is_offspring(unsigned int node) const {
    return std::find(children.begin(), children.end(), node) != children.end();
}

// person.cc:224-299.  NOTE the reference fills this cache from PedigreeParser::parse_line,
// i.e. while maternal_id/paternal_id still hold UNKNOWN_PARENT, so isfounder() is true for
// everyone and untyped non-founders receive the population prior.  Reproduced on purpose:
// the call site (parse_pedigree_file) runs before sanity_check() for the same effect.
void Person::populate_trait_prob_cache(const GeneticMap& map, bool sex_linked) {
    const unsigned M = (unsigned) genotypes.size();
    const bool xmale = sex_linked && ismale();
    genotypes_prob.assign((size_t) 4 * M, 0.0);
    for(unsigned i = 0; i < M; ++i) {
        double probs[4];
        for(int j = 0; j < 4; ++j) {
            const enum phased_trait pt = (enum phased_trait) j;
            const bool het = (pt == TRAIT_AU) || (pt == TRAIT_UA);
            const double marker_prob = map.get_prob(i, pt, xmale);
            const enum unphased_genotype g = (enum unphased_genotype) genotypes[i];
            double v;
            if(!isfounder()) {
                if(istyped()) {
                    switch(g) {
                        case HETERO:  v = het ? 1.0 : 0.0; break;
                        case HOMOZ_A: v = (pt == TRAIT_UU) ? 1.0 : 0.0; break;
                        case HOMOZ_B: v = (pt == TRAIT_AA) ? 1.0 : 0.0; break;
                        default:      v = 1.0; break;
                    }
                }
                else v = (xmale && het) ? 0.0 : 1.0;
            }
            else {
                if(istyped()) {
                    switch(g) {
                        case HETERO:  v = het ? marker_prob : 0.0; break;
                        case HOMOZ_A: v = (pt == TRAIT_UU) ? marker_prob : 0.0; break;
                        case HOMOZ_B: v = (pt == TRAIT_AA) ? marker_prob : 0.0; break;
                        default:      v = marker_prob; break;
                    }
                }
                else v = (xmale && het) ? 0.0 : marker_prob;
            }
            probs[j] = v;
        }
        double total = probs[0] + probs[1] + probs[2] + probs[3];
        for(int j = 0; j < 4; ++j) genotypes_prob[4 * i + j] = probs[j] / total;
    }
}

// person.cc:208-222
bool Person::safe_to_ignore_meiosis(const Pedigree& ped, enum parentage p, bool sex_linked) const {
    const Person* tmp = ped.get_by_index(p == MATERNAL ? maternal_id : paternal_id);
    if(!tmp->isfounder()) {
        if(sex_linked) return p == PATERNAL;
        return false;
    }
    return tmp->num_children() == 1;
}

// ---- pedigree ---------------------------------------------------------------------------------------

Person* Pedigree::get_by_name(const std::string& name) {
    for(unsigned int i = 0; i < members.size(); ++i) if(members[i].get_id() == name) return &members[i];
    return 0;
}

bool Pedigree::add(const Person& p) {
    if(exists(p.get_id())) return false;
    members.push_back(p);
    return true;
}

// genotype.cc:11-53
static bool genotype_compatible(enum unphased_genotype mother, enum unphased_genotype father,
                                enum unphased_genotype child, enum sex child_sex, bool sex_linked) {
    if(!sex_linked || child_sex == FEMALE) {
        switch(child) {
            case UNTYPED: return true;
            case HOMOZ_A: return (mother != HOMOZ_B) && (father != HOMOZ_B);
            case HOMOZ_B: return (mother != HOMOZ_A) && (father != HOMOZ_A);
            case HETERO:  return !((mother == HOMOZ_A) && (father == HOMOZ_A)) && !((mother == HOMOZ_B) && (father == HOMOZ_B));
        }
    }
    else {
        switch(child) {
            case UNTYPED: return true;
            case HETERO:  return false;
            case HOMOZ_A: return mother != HOMOZ_B;
            case HOMOZ_B: return mother != HOMOZ_A;
        }
    }
    abort();
}

bool Pedigree::mendelian_errors() const {
    for(unsigned int k = 0; k < members.size(); ++k) {
        const Person& c = members[k];
        if(c.isfounder_str()) continue;
        const Person& m = members[c.get_maternalid()];
        const Person& p = members[c.get_paternalid()];
        for(unsigned int i = 0; i < c.num_markers(); ++i) {
            if(!genotype_compatible(m.get_genotype(i), p.get_genotype(i), c.get_genotype(i), c.get_sex(), sex_linked)) {
                fprintf(stderr, "error: genotypes at loci number %d of person \"%s\" inconsistent with parents\n", i + 1, c.get_id().c_str());
                return true;
            }
        }
    }
    return false;
}

// pedigree.cc:230-287
int Pedigree::count_components() const {
    int components = 0;
    int total = (int) members.size();
    std::vector<int> visited(members.size(), 0);
    std::queue<unsigned int> q;
    do {
        for(unsigned int i = 0; i < members.size(); ++i) {
            if(visited[i] == 0) { q.push(i); visited[i] = 1; break; }
        }
        while(!q.empty()) {
            unsigned int cur = q.front();
            q.pop();
            const Person& p = members[cur];
            for(unsigned int i = 0; i < p.num_children(); ++i) {
                unsigned int c = p.get_child(i);
                if(visited[c] == 0) { visited[c] = 1; q.push(c); }
            }
            unsigned int m = p.get_maternalid(), f = p.get_paternalid();
            if(m != UNKNOWN_PARENT && visited[m] == 0) { visited[m] = 1; q.push(m); }
            if(f != UNKNOWN_PARENT && visited[f] == 0) { visited[f] = 1; q.push(f); }
            visited[cur] = 2;
            --total;
        }
        ++components;
    } while(total != 0);
    return components;
}

// pedigree.cc:58-90 and the helpers it calls
bool Pedigree::sanity_check() {
    // same number of markers
    for(size_t i = 1; i < members.size(); ++i) if(members[i].num_markers() != members[0].num_markers()) return false;

    // parents exist and have the right sex (pedigree.cc:106-152)
    bool error = false;
    for(unsigned int i = 0; i < members.size(); ++i) {
        Person* p = &members[i];
        Person* tmp;
        if(!p->mother_unknown()) {
            if((tmp = get_by_name(p->get_mother())) == 0) {
                fprintf(stderr, "error: %s, mother of %s does not exist\n", "_parental_relationship_errors", p->get_id().c_str());
                error = true;
            }
            else if(!tmp->isfemale()) {
                fprintf(stderr, "error: %s, mother of %s is not female\n", "_parental_relationship_errors", p->get_id().c_str());
                error = true;
            }
        }
        if(!p->father_unknown()) {
            if((tmp = get_by_name(p->get_father())) == 0) {
                fprintf(stderr, "error: %s, father of %s does not exist\n", "_parental_relationship_errors", p->get_id().c_str());
                error = true;
            }
            else if(!tmp->ismale()) {
                fprintf(stderr, "error: %s, father of %s is not male\n", "_parental_relationship_errors", p->get_id().c_str());
                error = true;
            }
        }
    }
    if(error) return false;

    // founders first: the reference uses std::sort with Person::operator< (pedigree.cc:154-162);
    // the same call on the same input order reproduces its (implementation-defined) permutation
    std::sort(members.begin(), members.end());
    for(unsigned int i = 0; i < members.size(); ++i) members[i].internal_id = i;

    // parent ids by name (pedigree.cc:164-193)
    for(unsigned int i = 0; i < members.size(); ++i) {
        Person* p = &members[i];
        if(p->mother_unknown()) p->maternal_id = UNKNOWN_PARENT;
        if(p->father_unknown()) p->paternal_id = UNKNOWN_PARENT;
        for(unsigned int j = 0; j < members.size(); ++j) {
            if(members[j].get_id() == p->get_mother()) p->maternal_id = j;
            if(members[j].get_id() == p->get_father()) p->paternal_id = j;
        }
    }

    // children / mates in pedigree order (person.cc:161-177)
    for(unsigned int k = 0; k < members.size(); ++k) {
        Person& me = members[k];
        me.children.clear();
        me.mates.clear();
        for(unsigned int i = 0; i < members.size(); ++i) {
            const Person& p = members[i];
            if(p.get_mother() == me.get_id()) {
                me.children.push_back(i);
                unsigned int mate = get_by_name(p.get_father())->get_internalid();
                if(std::find(me.mates.begin(), me.mates.end(), mate) == me.mates.end()) me.mates.push_back(mate);
            }
            if(p.get_father() == me.get_id()) {
                me.children.push_back(i);
                unsigned int mate = get_by_name(p.get_mother())->get_internalid();
                if(std::find(me.mates.begin(), me.mates.end(), mate) == me.mates.end()) me.mates.push_back(mate);
            }
        }
    }

    int components = count_components();
    if(components != 1) {
        fprintf(stderr, "error: %s, family %s is actually composed of %d distinct families\n", "sanity_check", id.c_str(), components);
        return false;
    }
    if(mendelian_errors()) return false;

    number_of_founders = number_of_leaves = 0;
    for(unsigned int i = 0; i < members.size(); ++i) {
        if(members[i].isfounder()) ++number_of_founders;
        if(members[i].isleaf()) ++number_of_leaves;
    }
    return true;
}

// ---- LINKAGE readers ---------------------------------------------------------------------------------

namespace {

// parser.h:60-103: '#' starts a comment, empty lines are skipped and not counted, and -- as in
// the reference's `while(!getline(f, line).eof())` -- a last line without a newline is not read
struct LineReader {
    std::ifstream f;
    int linenum;
    explicit LineReader(const std::string& fn) : f(fn.c_str()), linenum(0) {}
    bool ok() { return (bool) f; }
    bool next(std::string& line) {
        while(true) {
            if(std::getline(f, line).eof()) return false;
            std::string::size_type idx = line.find('#');
            if(idx != std::string::npos) line.erase(idx);
            if(line.empty()) continue;
            return true;
        }
    }
};

void tokenise(const std::string& s, std::vector<std::string>& tokens) {
    tokens.clear();
    std::istringstream ss(s);
    std::string t;
    while(ss >> t) tokens.push_back(t);
}

bool to_double(const std::string& s, double& out) {
    std::istringstream ss(s);
    return !(ss >> out).fail();
}

bool to_int(const std::string& s, int& out) {
    std::istringstream ss(s);
    return !(ss >> out).fail();
}

void get_doubles(const std::vector<std::string>& tokens, std::vector<double>& out) {
    out.clear();
    for(size_t i = 0; i < tokens.size(); ++i) {
        double v;
        if(to_double(tokens[i], v)) out.push_back(v);      // Mega2 appends "Haldane"/"Kosambi": skipped
    }
}

}  // namespace

// map_parser.cc:11-68
bool parse_map_file(const std::string& filename, GeneticMap& map) {
    LineReader r(filename);
    if(!r.ok()) { fprintf(stderr, "error: file not found: %s\n", filename.c_str()); return false; }
    std::string line;
    std::vector<std::string> tokens;
    bool noerror = true;
    while(r.next(line)) {
        tokenise(line, tokens);
        if(!tokens.empty()) {
            double gdist;
            if(tokens.size() < 3) {
                fprintf(stderr, "error: %s, line %d: not enough data fields specified (expected at least 3 (chromosome, genetic position, marker name), read %d)\n",
                        filename.c_str(), r.linenum + 1, int(tokens.size()));
                noerror = false;
            }
            else if(!to_double(tokens[1], gdist)) {
                fprintf(stderr, "error: %s, line %d: genetic distance is not a floating point number (read '%s')\n",
                        filename.c_str(), r.linenum + 1, tokens[1].c_str());
                noerror = false;
            }
            else if(gdist < 0.0) {
                fprintf(stderr, "error: %s, line %d: illegal genetic distance (%f)\n", filename.c_str(), r.linenum + 1, gdist);
                noerror = false;
            }
            else {
                gdist /= 100.0;                       // cM -> Morgans (map_parser.cc:62)
                map.add(Snp(tokens[2], gdist));
            }
        }
        r.linenum++;
    }
    return noerror && map.num_markers() > 1;
}

// linkage_parser.cc:18-439 (affection-status trait locus first, then numbered-allele SNPs)
bool parse_linkage_file(const std::string& filename, GeneticMap& map, DiseaseModel& dm) {
    LineReader r(filename);
    if(!r.ok()) { fprintf(stderr, "error: file not found: %s\n", filename.c_str()); return false; }
    std::string line;
    std::vector<std::string> tokens;
    std::vector<double> af;
    int number_of_loci = -1, marker_linenum = 0, recomb_linenum = 0, marker_code = -1, marker_alleles = -1;
    int markers_read[4] = {0, 0, 0, 0};
    double marker_freq = 0.0;
    const char* fn = filename.c_str();

    while(r.next(line)) {
        tokenise(line, tokens);
        const int ln = r.linenum;
        const int total_read = markers_read[0] + markers_read[1] + markers_read[2] + markers_read[3];
        if(ln == 0 || ln == 1) {
            if(tokens.size() != 4) {
                fprintf(stderr, "error: expected 4 fields on line %d of LINKAGE dat file (read %d)\n", ln + 1, int(tokens.size()));
                return false;
            }
            if(ln == 0) {
                int program_code;
                if(!to_int(tokens[0], number_of_loci)) {
                    fprintf(stderr, "error: %s, line 1: number of loci \"%s\" is not an integer\n", fn, tokens[0].c_str());
                    return false;
                }
                if(tokens[2] == "0") dm.set_sexlinked(false);
                else if(tokens[2] == "1") dm.set_sexlinked(true);
                else {
                    fprintf(stderr, "error: %s, line 1: sex-linked field \"%s\" must be either 0 or 1\n", fn, tokens[2].c_str());
                    return false;
                }
                if(!to_int(tokens[3], program_code)) {
                    fprintf(stderr, "error: %s, line 1: program code \"%s\" is not an integer\n", fn, tokens[3].c_str());
                    return false;
                }
            }
            else {
                const char* what[2] = {"mutation mode", "linkage disequilibrium"};
                const int idx[2] = {0, 3};
                for(int k = 0; k < 2; ++k) {
                    std::istringstream ss(tokens[idx[k]]);
                    bool b;
                    if((ss >> b).fail() || b) {
                        fprintf(stderr, "error: %s: %s should be set to 0\n", fn, what[k]);
                        return false;
                    }
                }
            }
        }
        else if(ln == 2) {
            if(int(tokens.size()) != number_of_loci) {
                fprintf(stderr, "error: %s: number of loci from line 1 (%d) and length of loci ordering from line 3 (%d) should match\n",
                        fn, number_of_loci, int(tokens.size()));
                return false;
            }
        }
        else if(total_read < number_of_loci) {
            if(marker_linenum == 0) {
                if(tokens.empty() || tokens[0].length() != 1 || !to_int(tokens[0], marker_code) || marker_code < 0 || marker_code > 3) {
                    fprintf(stderr, "error: %s, line %d: bad marker code %s, should be 0 - 3\n", fn, ln, tokens.empty() ? "" : tokens[0].c_str());
                    return false;
                }
            }
            if(marker_code == 0 || marker_code == 2) {
                fprintf(stderr, "Apologises: \"%s\" not supported\n", marker_code == 0 ? "quantitative variable" : "binary factor");
                return false;
            }
            if(marker_code == 1 && total_read != 0) {
                fprintf(stderr, "error: %s, line %d: trait marker must be the first marker in linkage file\n", fn, ln);
                return false;
            }
            if(marker_linenum == 0) {
                if(tokens.size() < 2 || !to_int(tokens[1], marker_alleles)) {
                    fprintf(stderr, "error: %s, line %d: first line of marker description should have been \"%s  N\" where N is the number of alleles\n",
                            fn, ln, tokens[0].c_str());
                    return false;
                }
                marker_linenum++;
            }
            else if(marker_linenum == 1) {
                get_doubles(tokens, af);
                if(marker_alleles != int(af.size())) {
                    fprintf(stderr, "error: %s, line %d: expected %d alleles, but read %d\n", fn, ln, marker_alleles, int(af.size()));
                    return false;
                }
                if(marker_alleles != 2) {
                    fprintf(stderr, "error: %s, line %d: this program is only designed to handle markers with 2 alleles\n", fn, ln);
                    return false;
                }
                if((af[0] + af[1]) < 0.99) {
                    fprintf(stderr, "error: %s, line %d: allele frequencies on this line do not sum to 1.0!\n", fn, ln + 1);
                    return false;
                }
                marker_freq = af[1];
                marker_linenum++;
                if(marker_code == 3) {
                    if(int(map.num_markers()) <= markers_read[3]) {
                        fprintf(stderr, "Error: more alleles in DAT file than MAP file!\nExiting...\n");
                        exit(EXIT_FAILURE);
                    }
                    map[markers_read[3]].set_minor_freq(marker_freq);
                    marker_linenum = 0;
                    markers_read[3]++;
                }
            }
            else if(marker_linenum == 2) {              // affection status: number of liability classes
                int classes;
                if(tokens.empty() || !to_int(tokens[0], classes)) {
                    fprintf(stderr, "error: %s, line %d: number of liability classes must be an integer (read \"%s\")\n",
                            fn, ln, tokens.empty() ? "" : tokens[0].c_str());
                    return false;
                }
                if(classes < 1) {
                    fprintf(stderr, "error: %s, line %d: must specify at least one liability class (%d specified)\n", fn, ln, classes);
                    return false;
                }
                if(classes != 1) fprintf(stderr, "Apologises: \"%s\" not supported\n", "more than one liability class");
                marker_linenum++;
            }
            else {                                      // penetrances
                get_doubles(tokens, af);
                if(int(af.size()) != 3) {
                    fprintf(stderr, "error: %s, line %d: liability classes should contain three numbers (read %d)\n", fn, ln, int(af.size()));
                    return false;
                }
                for(int i = 0; i < 3; ++i) dm.set_penetrance(af[i], (enum unphased_trait) i);
                dm.set_freq(marker_freq);
                marker_linenum = 0;
                markers_read[1]++;
            }
        }
        else {
            // recombination block: only the second line matters, and it is later overwritten by
            // the Haldane recomputation (genetic_map.cc:41-54)
            if(recomb_linenum == 1) {
                get_doubles(tokens, af);
                for(size_t i = 1; i < af.size(); ++i) map.add_theta(af[i]);
            }
            else if(recomb_linenum > 2) return false;
            recomb_linenum++;
        }
        r.linenum++;
    }
    if(marker_linenum != 0) { fprintf(stderr, "error: %s: unexpected EOF\n", fn); return false; }
    if(markers_read[1] != 1) { fprintf(stderr, "error: %s: did not read trait marker\n", fn); return false; }
    dm.finish_init();
    return true;
}

// pedigree_parser.cc:15-185
bool parse_pedigree_file(const std::string& filename, std::vector<Pedigree>& pedigrees, const DiseaseModel& dm,
                         const GeneticMap& map, bool ignore_genotypes) {
    LineReader r(filename);
    if(!r.ok()) { fprintf(stderr, "error: file not found: %s\n", filename.c_str()); return false; }
    std::string line;
    std::vector<std::string> tokens;
    bool noerror = true;
    const char* fn = filename.c_str();
    // note: Pedigree objects move when the vector grows, so look them up by index each line
    while(r.next(line)) {
        tokenise(line, tokens);
        const int ln = r.linenum;
        r.linenum++;
        if(tokens.empty()) continue;
        if(tokens.size() < 6) {
            fprintf(stderr, "error: %s, line %d: not enough fields, only %d found\n", fn, ln + 1, int(tokens.size()));
            noerror = false; continue;
        }
        if((tokens.size() % 2) != 0) {
            fprintf(stderr, "error: %s, line %d: contains an odd number of alleles\n", fn, ln + 1);
            noerror = false; continue;
        }
        enum sex s;
        enum affection a;
        if(tokens[4] == "0") s = UNSEXED; else if(tokens[4] == "1") s = MALE; else if(tokens[4] == "2") s = FEMALE;
        else {
            fprintf(stderr, "error: %s, line %d: bad sex \"%s\" (column 5)\n", fn, ln + 1, tokens[4].c_str());
            noerror = false; continue;
        }
        if(tokens[5] == "0") a = UNKNOWN_AFFECTION; else if(tokens[5] == "1") a = UNAFFECTED; else if(tokens[5] == "2") a = AFFECTED;
        else {
            fprintf(stderr, "error: %s, line %d: bad affection status \"%s\" (column 6)\n", fn, ln + 1, tokens[5].c_str());
            noerror = false; continue;
        }
        size_t pi = 0;
        for(; pi < pedigrees.size(); ++pi) if(pedigrees[pi].get_id() == tokens[0]) break;
        if(pi == pedigrees.size()) pedigrees.push_back(Pedigree(tokens[0], dm.is_sexlinked()));

        Person p(tokens[1], tokens[2], tokens[3], s, a, dm);
        bool bad = false;
        for(size_t i = 6; !ignore_genotypes && i + 1 < tokens.size(); i += 2) {      // pedigree_parser.cc:143
            const std::string& a1 = tokens[i];
            const std::string& a2 = tokens[i + 1];
            enum unphased_genotype g;
            if(a1 == "0" || a2 == "0") g = UNTYPED;
            else if(a1 == a2 && a1 == "1") g = HOMOZ_A;
            else if(a1 == a2 && a1 == "2") g = HOMOZ_B;
            else if((a1 == "1" && a2 == "2") || (a1 == "2" && a2 == "1")) g = HETERO;
            else {
                fprintf(stderr, "error: %s, line %d: error parsing genotype %d (columns %d and %d)\n",
                        fn, ln + 1, int((i - 6) / 2) + 1, int(i) + 1, int(i) + 2);
                bad = true;
                break;
            }
            p.add_genotype(g);
        }
        if(bad) { noerror = false; continue; }
        p.populate_trait_prob_cache(map, dm.is_sexlinked());      // before parent ids exist -- see the note above
        if(!pedigrees[pi].add(p)) noerror = false;
    }
    if(pedigrees.empty()) {
        fprintf(stderr, "error: %s contains no families\n", fn);
        return false;
    }
    for(size_t i = 0; i < pedigrees.size(); ++i) {
        if(!pedigrees[i].sanity_check()) {
            fprintf(stderr, "error: %s, pedigree %s contains relationship errors\n", fn, pedigrees[i].get_id().c_str());
            noerror = false;
        }
    }
    return noerror;
}

bool read_and_check_input(const std::string& pedfile, const std::string& mapfile, const std::string& datfile,
                          bool force_sex_linked, GeneticMap& map, DiseaseModel& dm, std::vector<Pedigree>& pedigrees) {
    if(!parse_map_file(mapfile, map)) return false;
    if(!parse_linkage_file(datfile, map, dm)) return false;
    if(force_sex_linked) dm.set_sexlinked(true);
    if(!map.sanity_check()) {
        fprintf(stderr, "Error: map data failed sanity check...\n");
        return false;
    }
    if(!parse_pedigree_file(pedfile, pedigrees, dm, map)) return false;
    for(size_t i = 0; i < pedigrees.size(); ++i) {
        if(map.num_markers() != pedigrees[i].num_markers()) {
            fprintf(stderr, "Error: different number of markers in \"%s\" pedigree (%d) versus map (%d)\nExiting...\n",
                    pedigrees[i].get_id().c_str(), pedigrees[i].num_markers(), map.num_markers());
            exit(EXIT_FAILURE);
        }
    }
    return true;
}

}  // namespace swiftlink
