// swift_main.cc -- the `swift` command line on top of the device path.  Same flags, input order,
// messages and output file as the reference's main.cc:185-557 / linkage_program.cc:26-171 for
// the linkage mode; every chain runs on the GPU (-g is implied, -X is allowed with it, which the
// reference refuses at main.cc:534-537): L-sampler and M-sampler sweeps, LOD scoring, the CODA trace.
// --elod runs the expected-LOD simulation on the device too (main.cc:570-577).
#include <getopt.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "swiftlink_host.h"

using namespace swiftlink;

static void usage(const char* prog) {
    struct mcmc_options d;
    fprintf(stderr,
"Usage: %s [OPTIONS] -p pedfile -m mapfile -d datfile\n"
"       %s [OPTIONS] -p pedfile --elod\n"
"\n"
"Input files:\n"
"  -p pedfile, --pedigree=pedfile\n"
"  -m mapfile, --map=mapfile\n"
"  -d datfile, --dat=datfile\n"
"\n"
"Output files:\n"
"  -o outfile, --output=outfile            (default = 'swiftlink.out')\n"
"\n"
"MCMC options:\n"
"  -i NUM,     --iterations=NUM            (default = %d)\n"
"  -b NUM,     --burnin=NUM                (default = %d)\n"
"  -s NUM,     --sequentialimputation=NUM  (default = %d)\n"
"  -x NUM,     --scoringperiod=NUM         (default = %d)\n"
"  -l FLOAT,   --lsamplerprobability=FLOAT (default = %.1f)\n"
"  -n NUM,     --lodscores=NUM             (default = %d)\n"
"  -R NUM,     --runs=NUM                  (default = %d)\n"
"\n"
"ELOD options:\n"
"  -e          --elod\n"
"  -f FLOAT    --frequency=FLOAT           (default = %.1e)\n"
"  -w FLOAT    --separation=FLOAT          (default = %.4f)\n"
"  -k FLOAT,FLOAT,FLOAT --penetrance=FLOAT,FLOAT,FLOAT(default = %.2f,%.2f,%.2f)\n"
"  -u NUM      --replicates=NUM            (default = %d)\n"
"\n"
"Metropolis-coupled MCMC options (commented out in the reference's main.cc:70-75):\n"
"  -M,         --mcmcmc\n"
"  -z NUM,     --chains=NUM                (default = %d)\n"
"  -y NUM,     --exchangeperiod=NUM        (default = %d)\n"
"  -t FLOAT,FLOAT,... --temperatures=FLOAT,FLOAT,...\n"
"  -j FILE,    --exchangefile=FILE\n"
"\n"
"Runtime options:\n"
"  -c NUM,     --cores=NUM                 replicate chains (-R) in flight on the device at once (default = %d)\n"
"  -g,         --gpu                       (always on)\n"
"  -D NUM,     --device=NUM                CUDA device (default = 0)\n"
"\n"
"Misc:\n"
"  -X,         --sexlinked\n"
"  -a,         --affectedonly\n"
"  -q NUM,     --peelseqiter=NUM           (default = %d)\n"
"  -S NUM,     --seed=NUM                  Philox seed (default = %llu)\n"
"  -T,         --trace\n"
"  -P PREFIX,  --traceprefix=PREFIX        (default = '%s')\n"
"  -v,         --verbose\n"
"  -h,         --help\n"
"\n",
    prog, prog, d.iterations, d.burnin, d.si_iterations, d.scoring_period, d.lsampler_prob, d.lodscores, d.mcmc_runs,
    d.elod_frequency, d.elod_marker_separation, d.elod_penetrance[0], d.elod_penetrance[1], d.elod_penetrance[2], d.elod_replicates,
    d.mc3_number_of_chains, d.mc3_exchange_period, d.thread_count, d.peelopt_iterations, (unsigned long long) d.seed, d.coda_prefix.c_str());
}

static bool str2int(int& out, const char* s) {
    char* end;
    long v = strtol(s, &end, 10);
    if(*s == '\0' || *end != '\0') return false;
    out = (int) v;
    return true;
}

int main(int argc, char** argv) {
    struct mcmc_options o;
    std::string pedfile, mapfile, datfile, outfile = "swiftlink.out";
    static struct option longopts[] = {
        {"affectedonly", no_argument, 0, 'a'}, {"burnin", required_argument, 0, 'b'}, {"cores", required_argument, 0, 'c'},
        {"dat", required_argument, 0, 'd'}, {"elod", no_argument, 0, 'e'}, {"gpu", no_argument, 0, 'g'},
        {"help", no_argument, 0, 'h'}, {"iterations", required_argument, 0, 'i'},
        {"lsamplerprobability", required_argument, 0, 'l'}, {"map", required_argument, 0, 'm'},
        {"lodscores", required_argument, 0, 'n'}, {"output", required_argument, 0, 'o'},
        {"pedigree", required_argument, 0, 'p'}, {"peelseqiter", required_argument, 0, 'q'},
        {"sequentialimputation", required_argument, 0, 's'}, {"verbose", no_argument, 0, 'v'},
        {"scoringperiod", required_argument, 0, 'x'}, {"sexlinked", no_argument, 0, 'X'},
        {"runs", required_argument, 0, 'R'}, {"trace", no_argument, 0, 'T'}, {"device", required_argument, 0, 'D'},
        {"seed", required_argument, 0, 'S'}, {"traceprefix", required_argument, 0, 'P'}, {"mcmcmc", no_argument, 0, 'M'},
        {"chains", required_argument, 0, 'z'}, {"exchangeperiod", required_argument, 0, 'y'},
        {"temperatures", required_argument, 0, 't'}, {"exchangefile", required_argument, 0, 'j'}, {"frequency", required_argument, 0, 'f'},
        {"separation", required_argument, 0, 'w'}, {"penetrance", required_argument, 0, 'k'}, {"replicates", required_argument, 0, 'u'},
        {0, 0, 0, 0}};
    int ch, tmp;
    while((ch = getopt_long(argc, argv, ":p:d:m:o:i:b:s:l:c:x:q:n:vhgeaXR:TD:S:P:Mz:y:t:j:f:w:k:u:", longopts, 0)) != -1) {
        switch(ch) {
            case 'p': pedfile = optarg; break;
            case 'm': mapfile = optarg; break;
            case 'd': datfile = optarg; break;
            case 'o': outfile = optarg; break;
            case 'v': o.verbose = true; break;
            case 'g': o.use_gpu = true; break;
            case 'a': o.affected_only = true; break;
            case 'X': o.sex_linked = true; break;
            case 'h': usage(argv[0]); return EXIT_SUCCESS;
            case 'e': o.elod = true; break;
            case 'f': o.elod_frequency = atof(optarg);
                      if(o.elod_frequency <= 0.0) { fprintf(stderr, "%s: ELOD trait frequency must be greater than zero ('%f' given)\n", argv[0], o.elod_frequency); return EXIT_FAILURE; }
                      break;
            case 'w': o.elod_marker_separation = atof(optarg);
                      if(o.elod_marker_separation <= 0.0) { fprintf(stderr, "%s: ELOD marker separation must be greater than zero ('%f' given)\n", argv[0], o.elod_marker_separation); return EXIT_FAILURE; }
                      break;
            case 'k': {
                double pen[3]; char extra;
                if(sscanf(optarg, "%lf,%lf,%lf%c", &pen[0], &pen[1], &pen[2], &extra) != 3) {
                    fprintf(stderr, "%s: penetrance requires 3 floats, e.g.: 0.0,0.0,1.0 ('%s' given)\n", argv[0], optarg); return EXIT_FAILURE;
                }
                for(int i = 0; i < 3; ++i) {
                    if(pen[i] < 0.0 || pen[i] > 1.0) { fprintf(stderr, "%s: penetraces must be comma delimited floats from 0.0 - 1.0 inclusive, e.g.: 0.0,0.0,1.0 ('%s' given)\n", argv[0], optarg); return EXIT_FAILURE; }
                    o.elod_penetrance[i] = pen[i];
                }
                break;
            }
            case 'T': o.coda_logging = true; break;
            case 'P': o.coda_prefix = optarg; break;
            case 'M': o.mc3 = true; break;
            case 'j': o.exchange_filename = optarg; break;
            case 't': {
                o.mc3_temperatures.clear();
                std::string str(optarg);
                size_t pos = 0;
                bool ok = !str.empty();
                while(ok && pos <= str.size()) {
                    size_t c = str.find(',', pos);
                    std::string tok = str.substr(pos, c == std::string::npos ? std::string::npos : c - pos);
                    char* end;
                    double v = strtod(tok.c_str(), &end);
                    if(tok.empty() || *end != '\0' || v < 0.0 || v > 1.0) ok = false;
                    else o.mc3_temperatures.push_back(v);
                    if(c == std::string::npos) break;
                    pos = c + 1;
                }
                if(!ok) { fprintf(stderr, "%s: temperatures must be comma delimited floats from 0.0 - 1.0 inclusive, e.g.: 1.0,0.9,0.8,0.7 ('%s' given)\n", argv[0], optarg); return EXIT_FAILURE; }
                break;
            }
            case 'l': o.lsampler_prob = atof(optarg);
                      if(o.lsampler_prob < 0.0 || o.lsampler_prob > 1.0) { fprintf(stderr, "%s: option '-l' requires a floating point argument between 0.0 and 1.0\n", argv[0]); return EXIT_FAILURE; }
                      break;
            case 'S': o.seed = strtoull(optarg, 0, 10); break;
            case 'i': case 'b': case 's': case 'x': case 'q': case 'n': case 'R': case 'c': case 'D': case 'z': case 'y': case 'u':
                if(!str2int(tmp, optarg) || tmp < 0) { fprintf(stderr, "%s: option '-%c' requires a non-negative integer argument ('%s' given)\n", argv[0], ch, optarg); return EXIT_FAILURE; }
                if(ch == 'i') o.iterations = tmp; else if(ch == 'b') o.burnin = tmp; else if(ch == 's') o.si_iterations = tmp;
                else if(ch == 'x') o.scoring_period = tmp; else if(ch == 'q') o.peelopt_iterations = tmp;
                else if(ch == 'n') o.lodscores = tmp; else if(ch == 'R') o.mcmc_runs = tmp; else if(ch == 'c') o.thread_count = tmp;
                else if(ch == 'z') o.mc3_number_of_chains = tmp; else if(ch == 'y') o.mc3_exchange_period = tmp;
                else if(ch == 'u') o.elod_replicates = tmp;
                else o.device = tmp;
                break;
            case ':': fprintf(stderr, "%s: option '-%c' requires an argument\n", argv[0], optopt); return EXIT_FAILURE;
            default:  fprintf(stderr, "%s: option '-%c' is invalid: ignored\n", argv[0], optopt); break;
        }
    }
    if(o.elod) {                                                      // main.cc:570-577, :609-610
        if(pedfile.empty()) { fprintf(stderr, "%s: --elod needs a pedigree file\n", argv[0]); return EXIT_FAILURE; }
        if(o.elod_replicates < 1) { fprintf(stderr, "%s: number of ELOD replicates must be greater than zero ('%d' given)\n", argv[0], o.elod_replicates); return EXIT_FAILURE; }
        Elod e(pedfile.c_str(), o);
        const double elod = e.run();
        fprintf(stderr, "\nELOD = %.3f\n", elod);
        return EXIT_SUCCESS;
    }
    if(pedfile.empty() || mapfile.empty() || datfile.empty()) {
        fprintf(stderr, "%s: the pedigree, map and dat files are all required\n", argv[0]);
        usage(argv[0]);
        return EXIT_FAILURE;
    }
    if(o.mc3_number_of_chains < 1 || o.mc3_exchange_period < 1) { fprintf(stderr, "%s: -z and -y must be at least 1\n", argv[0]); return EXIT_FAILURE; }
    if(!o.mc3_temperatures.empty() && (int) o.mc3_temperatures.size() != o.mc3_number_of_chains) {
        fprintf(stderr, "%s: %d temperature%s specified, but there are %d chain%s\n", argv[0], (int) o.mc3_temperatures.size(),
                o.mc3_temperatures.size() == 1 ? " was" : "s were", o.mc3_number_of_chains, o.mc3_number_of_chains == 1 ? "" : "s");   // main.cc:539-546
        return EXIT_FAILURE;
    }
    if(o.scoring_period < 1 || o.lodscores < 1 || o.mcmc_runs < 1) { fprintf(stderr, "%s: -x, -n and -R must be at least 1\n", argv[0]); return EXIT_FAILURE; }

    GeneticMap map(o.lodscores);
    DiseaseModel dm;
    std::vector<Pedigree> pedigrees;
    if(!read_and_check_input(pedfile, mapfile, datfile, o.sex_linked, map, dm, pedigrees)) {
        fprintf(stderr, "Exiting...\n");
        return EXIT_FAILURE;
    }
    o.sex_linked = dm.is_sexlinked();

    fprintf(stderr, "\nLinkage parameters:\n\tpenetrance = %.2f:%.2f:%.2f\n\ttrait freq = %.2e\n\tsex-linked = %s\n"
                    "\tburnin iterations = %d\n\tsampling iterations = %d\n\tsampling period = %d\n"
                    "\tlocus sampler prob = %.3f\n\tnumber of runs = %d\n\n",
            dm.get_penetrance(TRAIT_HOMO_U), dm.get_penetrance(TRAIT_HETERO), dm.get_penetrance(TRAIT_HOMO_A), dm.get_freq(),
            dm.is_sexlinked() ? "true" : "false", o.burnin, o.iterations, o.scoring_period, o.lsampler_prob, o.mcmc_runs);

    std::vector<LODscores*> all_scores;
    for(size_t pi = 0; pi < pedigrees.size(); ++pi) {
        Pedigree& p = pedigrees[pi];
        if(o.affected_only) {
            for(unsigned int i = 0; i < p.num_members(); ++i)
                if(!p.get_by_index(i)->isaffected()) p.get_by_index(i)->make_unknown_affection(dm);
        }
        // the peel sequence depends only on the pedigree: found once, shared by the replicates
        // (the reference repeats the search per replicate, linkage_program.cc:136-137)
        PeelSequenceGenerator psg(&p, &map, dm.is_sexlinked(), o.verbose, o.seed);
        psg.build_peel_sequence(o.peelopt_iterations);
        if(o.verbose) fprintf(stderr, "\n\n%s\n\n", psg.debug_string().c_str());

        LODscores* total = 0;
        if(o.mc3 && o.mc3_number_of_chains > 1) {
            for(int r = 0; r < o.mcmc_runs; ++r) {
                Mc3 ladder(&p, &map, &psg, o, r);                     // linkage_program.cc:169-170 (commented out upstream)
                LODscores* lod = ladder.run();
                if(!total) total = lod;
                else { total->merge_results(lod); delete lod; }
            }
        }
        else total = run_replicates(&p, &map, &psg, o, o.thread_count);
        all_scores.push_back(total);
    }
    bool ok = write_linkage_results(&map, outfile, all_scores, o.verbose);
    if(!ok) fprintf(stderr, "error: could not write output file '%s'\n", outfile.c_str());
    for(size_t i = 0; i < all_scores.size(); ++i) delete all_scores[i];
    return ok ? EXIT_SUCCESS : EXIT_FAILURE;
}
