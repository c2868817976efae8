/*
 * msampler_oracle.c -- CPU restatement of the reference's M-sampler (whole-chromosome Gibbs update
 * of one meiosis indicator) and of DescentGraph::get_likelihood.
 * TEST INFRASTRUCTURE ONLY: nothing under swiftlink_b200/ may include, link or call this.
 *
 * Parity status: PINNED.  tests/test_msampler_oracle.py checks, against the unmodified
 * reference compiled into oracle/_ref/libswiftref.so, on east / loop / xlinked / inbred:
 *   - every founder-allele-graph edge list and likelihood (bit-exact),
 *   - MeiosisSampler::reset / step: raw_matrix, fb_matrix after the backward pass and the sampled
 *     indicators, bit-exact, with the reference's own uniforms fed to orc_ms_step_stream,
 *   - DescentGraph::get_likelihood (bit-exact);
 * a sample of those outputs is frozen in tests/golden/<case>_ms.npz.
 *
 * Reference (file:line under /root/reference/src):
 *   FounderAlleleGraph4::reset / flip / propagate_fa_update   founder_allele_graph4.cc:548-598
 *   FounderAlleleGraph4::likelihood / combine_components      founder_allele_graph4.cc:34-546
 *   MeiosisSampler::reset / step / sample / graph_likelihood  meiosis_sampler.cc:17-203
 *   MarkovChain::_init (m_ordering), ::run (M-sweep)          markov_chain.cc:68-80, 342-349
 *   Person::safe_to_ignore_meiosis                            person.cc:208-222
 *   DescentGraph::get_likelihood and helpers                  descent_graph.cc:150-265
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "peel_oracle.h"
#include "philox.h"

#define LOG_ZERO (-DBL_MAX)
#define DEFAULT_COMPONENT (-1)

static int is_founder(const orc_problem* p, int i) { return p->mother[i] < 0 && p->father[i] < 0; }
static int dg_get(const orc_problem* p, const int* dg, int person, int locus, int parent) {
    return dg[(2 * p->N) * locus + 2 * person + parent];
}
static void dg_set(const orc_problem* p, int* dg, int person, int locus, int parent, int v) {
    dg[(2 * p->N) * locus + 2 * person + parent] = v;
}

/* meiosis_sampler.cc:41-72 (identical in descent_graph.cc:326-355) */
void orc_fag_sequence(const orc_problem* p, int* seq) {
    char* visited = (char*) calloc(p->N, 1);
    int n = 0, total = p->N;
    for(int i = 0; i < p->F; ++i) { seq[n++] = i; visited[i] = 1; total--; }
    while(total > 0) {
        for(int i = p->F; i < p->N; ++i) {
            if(visited[i]) continue;
            if(visited[p->mother[i]] && visited[p->father[i]]) { seq[n++] = i; visited[i] = 1; total--; }
        }
    }
    free(visited);
}

/* founder_allele_graph4.cc:548-572 */
void orc_fag_reset(const orc_problem* p, const int* dg, int locus, int* edge) {
    int* seq = (int*) malloc(sizeof(int) * p->N);
    orc_fag_sequence(p, seq);
    for(int i = 0; i < p->N; ++i) {
        int pid = seq[i], tmp = pid * 2;
        if(is_founder(p, pid)) {
            edge[tmp] = tmp;
            edge[tmp + 1] = tmp + 1;
        }
        else {
            edge[tmp] = edge[p->mother[pid] * 2 + dg_get(p, dg, pid, locus, 0)];
            edge[tmp + 1] = edge[p->father[pid] * 2 + dg_get(p, dg, pid, locus, 1)];
        }
    }
    free(seq);
}

/* founder_allele_graph4.cc:586-598 */
static void propagate(const orc_problem* p, const int* dg, int locus, int person, int allele, int allele_value,
                      int new_fa, int* edge) {
    if(dg_get(p, dg, person, locus, allele) != allele_value) return;
    int new_allele = (p->sex[person] == ORC_FEMALE) ? 0 : 1;
    for(int k = 0; k < p->N; ++k) {                     /* Person::children, pedigree order */
        if(is_founder(p, k)) continue;
        if(p->mother[k] == person || p->father[k] == person)
            propagate(p, dg, locus, k, new_allele, allele, new_fa, edge);
    }
    edge[person * 2 + allele] = new_fa;
}

/* founder_allele_graph4.cc:574-584 */
void orc_fag_flip(const orc_problem* p, const int* dg, int locus, int person, int parent, int* edge) {
    int allele_value = dg_get(p, dg, person, locus, parent);
    int par = parent == 0 ? p->mother[person] : p->father[person];
    int tmp = par * 2;
    int old_fa = edge[person * 2 + parent];
    int new_fa = (edge[tmp] == old_fa) ? edge[tmp + 1] : edge[tmp];
    propagate(p, dg, locus, person, parent, allele_value, new_fa, edge);
}

/* founder_allele_graph4.cc:427-455 */
static int fag_legal(int obs, int a1, int a2) {
    switch(obs) {
        case ORC_HETERO:  return a1 != a2;
        case ORC_HOMOZ_A: return a1 == ORC_HOMOZ_A && a2 == ORC_HOMOZ_A;
        case ORC_HOMOZ_B: return a1 == ORC_HOMOZ_B && a2 == ORC_HOMOZ_B;
    }
    abort();
}
static int fag_other(int obs, int a1) {
    switch(obs) {
        case ORC_HETERO:  return a1 == ORC_HOMOZ_A ? ORC_HOMOZ_B : ORC_HOMOZ_A;
        case ORC_HOMOZ_A: return a1 == ORC_HOMOZ_A ? ORC_HOMOZ_A : ORC_UNTYPED;
        case ORC_HOMOZ_B: return a1 == ORC_HOMOZ_B ? ORC_HOMOZ_B : ORC_UNTYPED;
    }
    abort();
}

typedef struct {
    int nfa;
    int* membership;      /* [2F] */
    int* fixed;           /* [2F] */
    int* active;          /* [2F] */
    int* assign[2];       /* [2][2F] */
    double* prob[2];      /* [2][2F] */
} fag_state;

static fag_state* fag_new(int nfa) {
    fag_state* s = (fag_state*) malloc(sizeof(fag_state));
    s->nfa = nfa;
    s->membership = (int*) malloc(sizeof(int) * nfa);
    s->fixed = (int*) malloc(sizeof(int) * nfa);
    s->active = (int*) malloc(sizeof(int) * nfa);
    for(int k = 0; k < 2; ++k) {
        s->assign[k] = (int*) calloc(nfa, sizeof(int));
        s->prob[k] = (double*) malloc(sizeof(double) * nfa);
        for(int i = 0; i < nfa; ++i) s->prob[k][i] = 1.0;
    }
    return s;
}
static void fag_free(fag_state* s) {
    free(s->membership); free(s->fixed); free(s->active);
    for(int k = 0; k < 2; ++k) { free(s->assign[k]); free(s->prob[k]); }
    free(s);
}

/* founder_allele_graph4.cc:504-546 */
static void combine(fag_state* s, int c1, int c2, int flip) {
    for(int i = 0; i < s->nfa; ++i) {
        if(s->membership[i] == c2) {
            s->membership[i] = c1;
            if(flip) { int t = s->assign[0][i]; s->assign[0][i] = s->assign[1][i]; s->assign[1][i] = t; }
        }
    }
    if(flip) { s->prob[0][c1] *= s->prob[1][c2]; s->prob[1][c1] *= s->prob[0][c2]; }
    else     { s->prob[0][c1] *= s->prob[0][c2]; s->prob[1][c1] *= s->prob[1][c2]; }
    s->active[c2] = 0;
}

/* one side already in a component `grp`, the other founder allele `fa_new` joins it
   (founder_allele_graph4.cc:286-332 and its mirror :334-384); fa_old is the member */
static int fag_join(fag_state* s, int g, int grp, int fa_old, int fa_new, double major, double minor) {
    int fx = s->fixed[grp];
    if(fx != -1) {
        int t0 = fag_other(g, s->assign[fx][fa_old]);
        if(t0 == ORC_UNTYPED) return 0;
        s->assign[fx][fa_new] = t0;
        s->prob[fx][grp] *= (t0 == ORC_HOMOZ_A) ? major : minor;
    }
    else {
        int t0 = fag_other(g, s->assign[0][fa_old]);
        int t1 = fag_other(g, s->assign[1][fa_old]);
        if(t0 != ORC_UNTYPED) {
            if(t1 != ORC_UNTYPED) {
                s->assign[0][fa_new] = t0;
                s->assign[1][fa_new] = t1;
                s->prob[0][grp] *= (t0 == ORC_HOMOZ_A) ? major : minor;
                s->prob[1][grp] *= (t1 == ORC_HOMOZ_A) ? major : minor;
            }
            else {
                s->assign[0][fa_new] = t0;
                s->prob[0][grp] *= (t0 == ORC_HOMOZ_A) ? major : minor;
                s->prob[1][grp] = 0.0;
                s->fixed[grp] = 0;
            }
        }
        else {
            if(t1 != ORC_UNTYPED) {
                s->assign[1][fa_new] = t1;
                s->prob[1][grp] *= (t1 == ORC_HOMOZ_A) ? major : minor;
                s->prob[0][grp] = 0.0;
                s->fixed[grp] = 1;
            }
            else return 0;
        }
    }
    s->membership[fa_new] = grp;
    return 1;
}

/* founder_allele_graph4.cc:34-424 */
static double fag_likelihood(const orc_problem* p, int locus, const int* edge, fag_state* s) {
    const double minor = p->minor[locus], major = 1.0 - minor;          /* genetic_map.h:42-47 */
#define FREQ(g) (((g) == ORC_HOMOZ_A) ? major : minor)
    int group_index = 0;
    for(int i = 0; i < s->nfa; ++i) { s->membership[i] = DEFAULT_COMPONENT; s->fixed[i] = -1; s->active[i] = 0; }

    for(int i = 0; i < p->N; ++i) {
        if(!p->typed[i]) continue;
        int g = p->genotypes[i * p->M + locus];
        if(g == ORC_UNTYPED) continue;
        int mat_fa = edge[2 * i], pat_fa = edge[2 * i + 1];

        if(mat_fa == pat_fa || (p->sex_linked && p->sex[i] == ORC_MALE)) {
            if(g == ORC_HETERO) return 0.0;
            int group1 = s->membership[mat_fa];
            if(group1 != DEFAULT_COMPONENT) {
                int fixed1 = s->fixed[group1];
                if(fixed1 != -1) {
                    if(g != s->assign[fixed1][mat_fa]) return 0.0;
                }
                else {
                    if(s->assign[0][mat_fa] == g)      { s->fixed[group1] = 0; s->prob[1][group1] = 0.0; }
                    else if(s->assign[1][mat_fa] == g) { s->fixed[group1] = 1; s->prob[0][group1] = 0.0; }
                    else return 0.0;
                }
            }
            else {
                s->membership[mat_fa] = group_index;
                s->fixed[group_index] = 0;
                s->active[group_index] = 1;
                s->assign[0][mat_fa] = g;
                s->prob[0][group_index] = FREQ(g);
                s->prob[1][group_index] = 0.0;
                ++group_index;
            }
            continue;
        }

        int group1 = s->membership[mat_fa], group2 = s->membership[pat_fa];
        if(group1 != DEFAULT_COMPONENT) {
            int fixed1 = s->fixed[group1];
            if(group2 != DEFAULT_COMPONENT) {
                if(group1 == group2) {
                    if(fixed1 != -1) {
                        if(!fag_legal(g, s->assign[fixed1][mat_fa], s->assign[fixed1][pat_fa])) return 0.0;
                    }
                    else {
                        int legal0 = fag_legal(g, s->assign[0][mat_fa], s->assign[0][pat_fa]);
                        int legal1 = fag_legal(g, s->assign[1][mat_fa], s->assign[1][pat_fa]);
                        if(legal0) {
                            if(!legal1) { s->fixed[group1] = 0; s->prob[1][group1] = 0.0; }
                        }
                        else {
                            if(legal1) { s->fixed[group1] = 1; s->prob[0][group1] = 0.0; }
                            else return 0.0;
                        }
                    }
                }
                else {
                    int fixed2 = s->fixed[group2];
                    if(fixed1 != -1) {
                        if(fixed2 != -1) {
                            if(!fag_legal(g, s->assign[fixed1][mat_fa], s->assign[fixed2][pat_fa])) return 0.0;
                        }
                        else {
                            int legal0 = fag_legal(g, s->assign[fixed1][mat_fa], s->assign[0][pat_fa]);
                            int legal1 = fag_legal(g, s->assign[fixed1][mat_fa], s->assign[1][pat_fa]);
                            if(legal0) fixed2 = 0;
                            else if(legal1) fixed2 = 1;
                            else return 0.0;
                        }
                    }
                    else if(fixed2 != -1) {
                        int legal0 = fag_legal(g, s->assign[0][mat_fa], s->assign[fixed2][pat_fa]);
                        int legal1 = fag_legal(g, s->assign[1][mat_fa], s->assign[fixed2][pat_fa]);
                        if(legal0) fixed1 = 0;
                        else if(legal1) fixed1 = 1;
                        else return 0.0;
                    }
                    else {
                        int legal0 = fag_legal(g, s->assign[0][mat_fa], s->assign[0][pat_fa]);
                        int legal1 = fag_legal(g, s->assign[1][mat_fa], s->assign[0][pat_fa]);
                        int legal2 = fag_legal(g, s->assign[0][mat_fa], s->assign[1][pat_fa]);
                        int legal3 = fag_legal(g, s->assign[1][mat_fa], s->assign[1][pat_fa]);
                        if(!(legal0 || legal1 || legal2 || legal3)) return 0.0;
                        if(legal0 && !(legal1 || legal2 || legal3))      { fixed1 = fixed2 = 0; }
                        else if(legal1 && !(legal0 || legal2 || legal3)) { fixed1 = 1; fixed2 = 0; }
                        else if(legal2 && !(legal0 || legal1 || legal3)) { fixed1 = 0; fixed2 = 1; }
                        else if(legal3 && !(legal0 || legal1 || legal2)) { fixed1 = fixed2 = 1; }
                        else if(legal0 && legal3 && !(legal1 || legal2)) { fixed1 = fixed2 = -1; }
                        else if(legal1 && legal2 && !(legal0 || legal3)) { fixed1 = fixed2 = -2; }
                        else                                              { fixed1 = fixed2 = -1; }
                    }

                    if(fixed1 != fixed2) {
                        s->fixed[group1] = fixed1;
                        s->prob[1 - fixed1][group1] = 0.0;
                        s->prob[1 - fixed2][group2] = 0.0;
                        combine(s, group1, group2, 1);
                    }
                    else {
                        if(fixed1 == -2) {
                            s->fixed[group1] = -1;
                            combine(s, group1, group2, 1);
                        }
                        else {
                            s->fixed[group1] = fixed1;
                            if(fixed1 != -1) {
                                s->prob[1 - fixed1][group1] = 0.0;
                                s->prob[1 - fixed2][group2] = 0.0;
                            }
                            combine(s, group1, group2, 0);
                        }
                    }
                }
            }
            else {
                if(!fag_join(s, g, group1, mat_fa, pat_fa, major, minor)) return 0.0;
            }
        }
        else if(group2 != DEFAULT_COMPONENT) {
            if(!fag_join(s, g, group2, pat_fa, mat_fa, major, minor)) return 0.0;
        }
        else {
            if(g == ORC_HETERO) {
                s->assign[0][mat_fa] = s->assign[1][pat_fa] = ORC_HOMOZ_A;
                s->assign[1][mat_fa] = s->assign[0][pat_fa] = ORC_HOMOZ_B;
                s->prob[0][group_index] = s->prob[1][group_index] = major * minor;
                s->fixed[group_index] = -1;
            }
            else {
                s->assign[0][mat_fa] = s->assign[0][pat_fa] = g;
                s->prob[0][group_index] = FREQ(g) * FREQ(g);
                s->prob[1][group_index] = 0.0;
                s->fixed[group_index] = 0;
            }
            s->membership[mat_fa] = s->membership[pat_fa] = group_index;
            s->active[group_index] = 1;
            ++group_index;
        }
    }

    double ret = 1.0;
    for(int i = 0; i < group_index; ++i) {
        if(s->active[i]) {
            int fx = s->fixed[i];
            if(fx != -1) ret *= s->prob[fx][i];
            else ret *= (s->prob[0][i] + s->prob[1][i]);
        }
    }
    return ret;
#undef FREQ
}

double orc_fag_likelihood(const orc_problem* p, int locus, const int* edge) {
    fag_state* s = fag_new(2 * p->F);
    double r = fag_likelihood(p, locus, edge, s);
    fag_free(s);
    return r;
}

/* ---- MeiosisSampler -------------------------------------------------------------------- */

struct orc_msampler {
    const orc_problem* p;
    int* edges;            /* [M][2N]  f4[locus].edge_list */
    double* raw;           /* [M][2] */
    double* fb;            /* [M][2] */
    double* fwd;           /* [M][2] snapshot of fb after the forward pass */
    fag_state* scratch;
    int last_parameter;
};

orc_msampler* orc_ms_create(const orc_problem* p) {
    orc_msampler* m = (orc_msampler*) calloc(1, sizeof(orc_msampler));
    m->p = p;
    m->edges = (int*) calloc((size_t) p->M * 2 * p->N, sizeof(int));
    m->raw = (double*) calloc((size_t) p->M * 2, sizeof(double));
    m->fb = (double*) calloc((size_t) p->M * 2, sizeof(double));
    m->fwd = (double*) calloc((size_t) p->M * 2, sizeof(double));
    m->scratch = fag_new(2 * p->F);
    return m;
}

void orc_ms_destroy(orc_msampler* m) {
    if(!m) return;
    free(m->edges); free(m->raw); free(m->fb); free(m->fwd);
    fag_free(m->scratch);
    free(m);
}

/* meiosis_sampler.cc:74-102 */
static double graph_likelihood(orc_msampler* m, const int* dg, int person, int locus, int parent, int value) {
    const orc_problem* p = m->p;
    int* edge = m->edges + (size_t) locus * 2 * p->N;
    int flip = dg_get(p, dg, person, locus, parent) != value;
    if(flip) orc_fag_flip(p, dg, locus, person, parent, edge);
    double lik = fag_likelihood(p, locus, edge, m->scratch);
    if(flip) orc_fag_flip(p, dg, locus, person, parent, edge);
    return lik;
}

/* meiosis_sampler.cc:17-39; returns 0, or 1 + locus of an illegal graph */
int orc_ms_reset(orc_msampler* m, const int* dg, int parameter) {
    const orc_problem* p = m->p;
    int person = p->F + parameter / 2, par = parameter % 2;
    for(int i = 0; i < p->M; ++i) orc_fag_reset(p, dg, i, m->edges + (size_t) i * 2 * p->N);
    for(int i = 0; i < p->M; ++i) {
        int meiosis = dg_get(p, dg, person, i, par);
        m->raw[2 * i + meiosis] = graph_likelihood(m, dg, person, i, par, meiosis);
        if(m->raw[2 * i + meiosis] == 0.0) return 1 + i;
    }
    m->last_parameter = parameter;
    return 0;
}

typedef double (*draw_fn)(void* ctx, int locus);

/* meiosis_sampler.cc:193-203 */
static int ms_sample(orc_msampler* m, int locus, draw_fn draw, void* ctx) {
    int index = locus * 2;
    if(m->fb[index] == 0.0) return 1;
    if(m->fb[index + 1] == 0.0) return 0;
    return (draw(ctx, locus) < (m->fb[index] / (m->fb[index] + m->fb[index + 1]))) ? 0 : 1;
}

/* meiosis_sampler.cc:104-191 */
static int ms_step(orc_msampler* m, int* dg, int parameter, draw_fn draw, void* ctx) {
    const orc_problem* p = m->p;
    int person = p->F + parameter / 2, par = parameter % 2;
    int last_id = p->F + m->last_parameter / 2, last_p = m->last_parameter % 2;
    const int M = p->M;

    for(int i = 0; i < M; ++i) {
        int index = i * 2;
        int tmp = dg_get(p, dg, person, i, par);
        int tmp2 = dg_get(p, dg, last_id, i, last_p);
        m->raw[index + tmp] = m->raw[index + tmp2];
        m->raw[index + (1 - tmp)] = graph_likelihood(m, dg, person, i, par, 1 - tmp);
        if(m->raw[index] == 0.0 && m->raw[index + 1] == 0.0) return 1 + i;
    }

    double total = m->raw[0] + m->raw[1];
    m->fb[0] = m->raw[0] / total;
    m->fb[1] = m->raw[1] / total;
    for(int i = 1; i < M; ++i) {
        int index = i * 2;
        for(int j = 0; j < 2; ++j) {
            m->fb[index + j] = m->raw[index + j] *
                ((m->fb[(i - 1) * 2 + (1 - j)] * p->theta[i - 1]) +
                 (m->fb[(i - 1) * 2 + j] * (1.0 - p->theta[i - 1])));
        }
        double t = m->fb[index] + m->fb[index + 1];
        m->fb[index] /= t;
        m->fb[index + 1] /= t;
    }
    memcpy(m->fwd, m->fb, sizeof(double) * 2 * M);

    int i = M - 1;
    int orig = dg_get(p, dg, person, i, par);
    int samp = ms_sample(m, i, draw, ctx);
    if(orig != samp) {
        orc_fag_flip(p, dg, i, person, par, m->edges + (size_t) i * 2 * p->N);
        dg_set(p, dg, person, i, par, samp);
    }
    while(--i >= 0) {
        int index = i * 2;
        for(int j = 0; j < 2; ++j) {
            double next = (dg_get(p, dg, person, i + 1, par) != j) ? p->theta[i] : (1.0 - p->theta[i]);
            m->fb[index + j] *= next;
        }
        orig = dg_get(p, dg, person, i, par);
        samp = ms_sample(m, i, draw, ctx);
        if(orig != samp) {
            orc_fag_flip(p, dg, i, person, par, m->edges + (size_t) i * 2 * p->N);
            dg_set(p, dg, person, i, par, samp);
        }
    }
    m->last_parameter = parameter;
    return 0;
}

typedef struct { const double* us; int n, used; } stream_ctx;
static double draw_stream(void* c, int locus) {
    stream_ctx* s = (stream_ctx*) c;
    (void) locus;
    if(s->used >= s->n) abort();
    return s->us[s->used++];
}

typedef struct { uint64_t seed; uint32_t chain; uint64_t iteration; uint32_t slot; } philox_ctx;
static double draw_philox(void* c, int locus) {
    philox_ctx* s = (philox_ctx*) c;
    return orc_uniform(s->seed, s->chain, s->iteration, (uint32_t) locus, s->slot);
}

/* step with the caller's uniforms, consumed in the reference's order (last locus first, one
   per locus whose two weights are both non-zero); *used receives how many were taken */
int orc_ms_step_stream(orc_msampler* m, int* dg, int parameter, const double* us, int n, int* used) {
    stream_ctx c = { us, n, 0 };
    int rc = ms_step(m, dg, parameter, draw_stream, &c);
    if(used) *used = c.used;
    return rc;
}

/* step with the shared Philox schedule: the draw of locus l is
   uniform(seed, chain, iteration, l, ORC_SLOT_MEIOSIS + parameter) */
int orc_ms_step(orc_msampler* m, int* dg, int parameter, uint64_t seed, uint32_t chain, uint64_t iteration) {
    philox_ctx c = { seed, chain, iteration, ORC_SLOT_MEIOSIS + (uint32_t) parameter };
    return ms_step(m, dg, parameter, draw_philox, &c);
}

void orc_ms_state(const orc_msampler* m, double* raw, double* fwd, double* fb, int* edges) {
    const orc_problem* p = m->p;
    if(raw) memcpy(raw, m->raw, sizeof(double) * 2 * p->M);
    if(fwd) memcpy(fwd, m->fwd, sizeof(double) * 2 * p->M);
    if(fb)  memcpy(fb, m->fb, sizeof(double) * 2 * p->M);
    if(edges) memcpy(edges, m->edges, sizeof(int) * (size_t) p->M * 2 * p->N);
}

/* markov_chain.cc:68-80 + person.cc:208-222: the meioses an M-sweep visits, pedigree order */
int orc_ms_ordering(const orc_problem* p, int* out) {
    int n = 0;
    int* nchild = (int*) calloc(p->N, sizeof(int));
    for(int i = p->F; i < p->N; ++i) { nchild[p->mother[i]]++; nchild[p->father[i]]++; }
    for(int i = 0; i < 2 * (p->N - p->F); ++i) {
        int person = p->F + i / 2, par = i % 2;
        int parent = par == 0 ? p->mother[person] : p->father[person];
        int ignore;
        if(!is_founder(p, parent)) ignore = p->sex_linked ? (par == 1) : 0;
        else ignore = nchild[parent] == 1;
        if(!ignore) out[n++] = i;
    }
    free(nchild);
    return n;
}

/* Fisher-Yates with the shared Philox schedule (the reference's random_shuffle draws from libc
   rand(), markov_chain.cc:343 -- unpinned, replaced like every other draw):
   for i = n-1 .. 1: j = floor(u_i * (i + 1)), u_i = uniform(seed, chain, iteration, i, ORC_SLOT_MSHUFFLE) */
void orc_ms_shuffle(int* v, int n, uint64_t seed, uint32_t chain, uint64_t iteration) {
    for(int i = n - 1; i > 0; --i) {
        int j = (int)(orc_uniform(seed, chain, iteration, (uint32_t) i, ORC_SLOT_MSHUFFLE) * (double)(i + 1));
        if(j > i) j = i;
        int t = v[i]; v[i] = v[j]; v[j] = t;
    }
}

/* One M-sweep (markov_chain.cc:342-349).  Returns 0, or 1 + locus if a graph was illegal. */
int orc_ms_sweep(const orc_problem* p, int* dg, uint64_t seed, uint32_t chain, uint64_t iteration) {
    int* order = (int*) malloc(sizeof(int) * 2 * (p->N - p->F) + sizeof(int));
    int n = orc_ms_ordering(p, order);
    int rc = 0;
    if(n > 0) {
        orc_msampler* m = orc_ms_create(p);
        orc_ms_shuffle(order, n, seed, chain, iteration);
        rc = orc_ms_reset(m, dg, order[0]);
        for(int j = 0; j < n && rc == 0; ++j) rc = orc_ms_step(m, dg, order[j], seed, chain, iteration);
        orc_ms_destroy(m);
    }
    free(order);
    return rc;
}

/* ---- DescentGraph::get_likelihood (descent_graph.cc:150-265) ----------------------------- */

/* sum over loci of ln(founder allele graph likelihood); LOG_ZERO if a locus is illegal */
double orc_dg_sum_prior_prob(const orc_problem* p, const int* dg) {
    int* edge = (int*) malloc(sizeof(int) * 2 * p->N);
    fag_state* s = fag_new(2 * p->F);
    double ret = 0.0;
    for(int i = 0; i < p->M; ++i) {
        orc_fag_reset(p, dg, i, edge);
        double t = fag_likelihood(p, i, edge, s);
        if(t == 0.0) { ret = LOG_ZERO; break; }
        ret += log(t);
    }
    fag_free(s);
    free(edge);
    return ret;
}

double orc_dg_likelihood(const orc_problem* p, const int* dg) {
    double rec = 0.0;
    for(int i = 0; i < p->M - 1; ++i) rec += orc_recombination_prob(p, dg, i);
    double trans = orc_marker_transmission(p) + rec;
    double prior = orc_dg_sum_prior_prob(p, dg);
    return (trans == LOG_ZERO || prior == LOG_ZERO) ? LOG_ZERO : trans + prior;    /* logarithms.cc:25-27 */
}
