/*
 * peel_oracle.c -- plain-C restatement of the reference's peeling R-function path.
 * TEST INFRASTRUCTURE ONLY (see peel_oracle.h for the parity status and the import rule).
 *
 * Each function names the reference lines (under /root/reference/src) it follows.  The
 * arithmetic is written in the same operation order as the reference so that results agree to
 * the last bit wherever the reference itself is deterministic; compile with -ffp-contract=off.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

#include "peel_oracle.h"
#include "philox.h"

#define MATERNAL 0
#define PATERNAL 1
#define LOG_ZERO (-DBL_MAX)                                     /* logarithms.h:8-9 */

/* ---- small helpers ---------------------------------------------------------------- */

static long ipow4(int c) { return 1L << (2 * c); }

long orc_matrix_doubles(const orc_problem* p) {
    long n = 0;
    for(int i = 0; i < p->nops; ++i) n += ipow4(p->ops[i].ncut);
    return n;
}

long orc_presum_doubles(const orc_problem* p) {
    long n = 0;
    for(int i = 0; i < p->nops; ++i) n += ipow4(p->ops[i].ncut + 1);
    return n;
}

/* genotype.cc:91-102 */
static int mask_from_trait(int v) {
    switch(v) {
        case ORC_UU: return 8;   /* AA */
        case ORC_AU: return 2;   /* BA */
        case ORC_UA: return 4;   /* AB */
        case ORC_AA: return 1;   /* BB */
    }
    abort();
}

/* elimination.cc:393-395 */
int orc_is_legal(const orc_problem* p, int person, int locus, int value) {
    return (p->elim[(long) locus * p->N + person] & mask_from_trait(value)) != 0;
}

/* added to the locus in every Philox key: ELOD replicate r draws as if its three loci were rows
   3r .. 3r+2 of one long graph (orc_elod_replicate) */
static long g_locus_key_offset = 0;

static int dg_get(const orc_problem* p, const int* dg, int person, int locus, int parent) {
    return dg[((long) locus * p->N + person) * 2 + parent];      /* descent_graph.h:35-37 */
}

static int is_founder(const orc_problem* p, int i) { return p->mother[i] < 0 && p->father[i] < 0; }

/* rfunction.h:81-107 */
static int affected_trait(int pt, int allele) {
    switch(pt) {
        case ORC_UU: return 0;
        case ORC_AU: return allele == 0;
        case ORC_UA: return allele == 1;
        case ORC_AA: return 1;
    }
    abort();
}

/* rfunction.cc:94-113 */
static int phased_trait(const orc_problem* p, int m, int f, int mat_allele, int pat_allele, int child_sex) {
    int ma = affected_trait(m, mat_allele);
    int pa = affected_trait(f, pat_allele);
    if(p->sex_linked && child_sex == ORC_MALE) return ma ? ORC_AA : ORC_UU;
    if(ma) return pa ? ORC_AA : ORC_AU;
    return pa ? ORC_UA : ORC_UU;
}

/* ---- index tables (peel_sequence_generator.cc:84-186) ------------------------------ */

int orc_op_indices(const orc_problem* p, int opi, int which, int locus, int* out, int cap) {
    const orc_op* op = &p->ops[opi];
    int c = op->ncut;
    int ndig = (which == 2) ? c + 1 : c;
    long total = ipow4(ndig);
    int n = 0;
    for(long idx = 0; idx < total; ++idx) {
        int valid = 1;
        for(int d = 0; d < ndig && valid; ++d) {
            int person = (d < c) ? op->cutset[d] : op->peelnode;
            int v = (int)((idx >> (2 * d)) & 3);
            if(which == 0) {
                if(p->disease_prob[person * 4 + v] == 0.0) valid = 0;      /* :162-179 */
            }
            else {
                if(!orc_is_legal(p, person, locus, v)) valid = 0;          /* :109-159 */
            }
        }
        if(valid) {
            if(n < cap) out[n] = (int) idx;
            ++n;
        }
    }
    return n;
}

/* ---- marker genotype prior (person.cc:224-299) ------------------------------------- */

static double genotype_probability(int isfounder, int typed, int g, int xmale, int pt, double marker_prob) {
    int het = (pt == ORC_AU) || (pt == ORC_UA);
    if(!isfounder) {
        if(typed) {
            switch(g) {
                case ORC_HETERO:  return het ? 1.0 : 0.0;
                case ORC_HOMOZ_A: return (pt == ORC_UU) ? 1.0 : 0.0;
                case ORC_HOMOZ_B: return (pt == ORC_AA) ? 1.0 : 0.0;
                default:          return 1.0;
            }
        }
        if(xmale) return het ? 0.0 : 1.0;
        return 1.0;
    }
    if(typed) {
        switch(g) {
            case ORC_HETERO:  return het ? marker_prob : 0.0;
            case ORC_HOMOZ_A: return (pt == ORC_UU) ? marker_prob : 0.0;
            case ORC_HOMOZ_B: return (pt == ORC_AA) ? marker_prob : 0.0;
            default:          return marker_prob;
        }
    }
    if(xmale) return het ? 0.0 : marker_prob;
    return marker_prob;
}

void orc_marker_prob(int isfounder, int typed, int genotype, int xmale, const double mapprob[4], double out[4]) {
    double total;
    for(int j = 0; j < 4; ++j) out[j] = genotype_probability(isfounder, typed, genotype, xmale, j, mapprob[j]);
    total = out[0] + out[1] + out[2] + out[3];
    for(int j = 0; j < 4; ++j) out[j] /= total;
}

/* ---- evaluation context ------------------------------------------------------------ */

typedef struct {
    const orc_problem* p;
    const int* dg;
    int locus;
    int trait_mode;              /* 0 = SamplerRfunction, 1 = TraitRfunction */
    int ignore_left, ignore_right;
    double theta, antitheta, theta2, antitheta2;
    double** mat;                /* per op, 4^c   */
    double** pre;                /* per op, 4^(c+1) (sampler only) */
    int* assign;                 /* [N] scratch, the reference's indices[pmatrix_index] row */
    double* tmat;                /* [ORC_MAXC][64] transmission tables of the current op */
} orc_ctx;

static orc_ctx* ctx_new(const orc_problem* p, int want_presum) {
    orc_ctx* c = (orc_ctx*) calloc(1, sizeof(orc_ctx));
    c->p = p;
    c->mat = (double**) calloc(p->nops, sizeof(double*));
    c->pre = (double**) calloc(p->nops, sizeof(double*));
    for(int i = 0; i < p->nops; ++i) {
        c->mat[i] = (double*) calloc(ipow4(p->ops[i].ncut), sizeof(double));
        if(want_presum) c->pre[i] = (double*) calloc(ipow4(p->ops[i].ncut + 1), sizeof(double));
    }
    c->assign = (int*) malloc(sizeof(int) * p->N);
    c->tmat = (double*) malloc(sizeof(double) * 64 * ORC_MAXC);
    return c;
}

static void ctx_free(orc_ctx* c) {
    for(int i = 0; i < c->p->nops; ++i) { free(c->mat[i]); free(c->pre[i]); }
    free(c->mat); free(c->pre); free(c->assign); free(c->tmat); free(c);
}

/* peel_matrix.h:37-45 applied to a previous function's keys */
static double prev_get(const orc_ctx* c, int opi) {
    const orc_op* op = &c->p->ops[opi];
    long idx = 0;
    for(int i = 0; i < op->ncut; ++i) idx += (long) c->assign[op->cutset[i]] << (2 * i);
    return c->mat[opi][idx];
}

/* ---- sampler R-function ------------------------------------------------------------ */

/* sampler_rfunction.h:84-137 (set_locus / set_locus_minimal; the stale end thetas of the
   minimal variant are never read, so one routine serves both) */
static void sampler_set_locus(orc_ctx* c, int locus, int ignore_left, int ignore_right) {
    const orc_problem* p = c->p;
    c->locus = locus;
    c->ignore_left = ignore_left;
    c->ignore_right = ignore_right;
    c->theta = c->theta2 = c->antitheta = c->antitheta2 = 1.0;
    if(locus != 0 && !ignore_left) {
        c->theta2 = p->theta[locus - 1];
        c->antitheta2 = 1.0 - p->theta[locus - 1];              /* genetic_map.h:151-152 */
    }
    if(locus != p->M - 1 && !ignore_right) {
        c->theta = p->theta[locus];
        c->antitheta = 1.0 - p->theta[locus];
    }
}

/* sampler_rfunction.cc:102-157 */
static void recombination_distribution(const orc_ctx* c, int person, int parent_trait, int parent, double dist[2]) {
    const orc_problem* p = c->p;
    double tmp0 = 0.5, tmp1 = 0.5, total;
    switch(parent_trait) {
        case ORC_UU: dist[0] = 1.0; dist[1] = 0.0; return;
        case ORC_AA: dist[0] = 0.0; dist[1] = 1.0; return;
        default:
            if(p->sex_linked && parent == PATERNAL) { dist[0] = 0.0; dist[1] = 0.0; return; }
    }
    if(c->locus != 0) {
        int cross = dg_get(p, c->dg, person, c->locus - 1, parent) != 0;
        tmp0 *= cross ? c->theta2 : c->antitheta2;
        tmp1 *= cross ? c->antitheta2 : c->theta2;
    }
    if(c->locus != p->M - 1) {
        int cross = dg_get(p, c->dg, person, c->locus + 1, parent) != 0;
        tmp0 *= cross ? c->theta : c->antitheta;
        tmp1 *= cross ? c->antitheta : c->theta;
    }
    total = tmp0 + tmp1;
    if(parent_trait == ORC_AU) dist[0] = tmp1 / total;
    else                       dist[0] = tmp0 / total;
    dist[1] = 1.0 - dist[0];
}

/* sampler_rfunction.cc:350-419 */
static void transmission_matrix(const orc_ctx* c, int kid, double* t) {
    const orc_problem* p = c->p;
    double md[2], pd[2];
    int male = p->sex[kid] == ORC_MALE;
    for(int i = 0; i < 4; ++i) {
        recombination_distribution(c, kid, i, MATERNAL, md);
        for(int j = 0; j < 4; ++j) {
            double* row = t + 16 * i + 4 * j;
            recombination_distribution(c, kid, j, PATERNAL, pd);
            if(p->sex_linked && male) {
                if(j == ORC_AU || j == ORC_UA) {
                    row[ORC_UU] = row[ORC_AU] = row[ORC_UA] = row[ORC_AA] = 0.0;
                }
                else {
                    row[ORC_UU] = md[0];
                    row[ORC_AU] = 0.0;
                    row[ORC_UA] = 0.0;
                    row[ORC_AA] = md[1];
                }
            }
            else {
                row[ORC_UU] = md[0] * pd[0];
                row[ORC_AU] = md[1] * pd[0];
                row[ORC_UA] = md[0] * pd[1];
                row[ORC_AA] = md[1] * pd[1];
            }
        }
    }
}

/* rfunction.cc:116-142, sampler_rfunction.cc:190-233 and :236-287 in one routine; the three
   differ only in the transmission factor */
static void sampler_element(orc_ctx* c, int opi, long idx, const double trait_cache[4]) {
    const orc_problem* p = c->p;
    const orc_op* op = &p->ops[opi];
    long offset = ipow4(op->ncut);
    double total = 0.0;
    int mat_trait = 0, pat_trait = 0;

    if(op->type == ORC_CHILD_PEEL) {
        mat_trait = c->assign[p->mother[op->peelnode]];
        pat_trait = c->assign[p->father[op->peelnode]];
    }

    for(int g = 0; g < 4; ++g) {
        double tmp = trait_cache[g];
        c->assign[op->peelnode] = g;
        if(tmp == 0.0) continue;

        if(op->type == ORC_CHILD_PEEL) {
            tmp *= c->tmat[16 * mat_trait + 4 * pat_trait + g];
        }

        for(int j = 0; j < op->nprev; ++j) tmp *= prev_get(c, op->prev[j]);

        if(op->type == ORC_PARENT_PEEL) {
            double child_prob = 1.0;
            int m = g, f = g;
            for(int k = 0; k < op->nchild; ++k) {
                int kid = op->children[k];
                int kid_trait = c->assign[kid];
                if(p->mother[kid] == op->peelnode) f = c->assign[p->father[kid]];
                else                               m = c->assign[p->mother[kid]];
                child_prob *= c->tmat[64 * k + 16 * m + 4 * f + kid_trait];
            }
            tmp *= child_prob;
        }

        c->pre[opi][idx + offset * g] = tmp;
        total += tmp;
    }
    c->mat[opi][idx] = total;
}

/* rfunction.cc:172-198 (offset == 0 branch) with sampler_rfunction.cc:328-341 */
static void sampler_evaluate(orc_ctx* c, int opi) {
    const orc_problem* p = c->p;
    const orc_op* op = &p->ops[opi];
    long size = ipow4(op->ncut);
    double trait_cache[4];

    memset(c->mat[opi], 0, sizeof(double) * size);                /* pmatrix.reset() */
    memset(c->pre[opi], 0, sizeof(double) * size * 4);            /* pmatrix_presum.reset() */
    for(int g = 0; g < 4; ++g)
        trait_cache[g] = p->marker_prob[((long) op->peelnode * p->M + c->locus) * 4 + g];

    if(op->type == ORC_PARENT_PEEL) {
        for(int k = 0; k < op->nchild; ++k) transmission_matrix(c, op->children[k], c->tmat + 64 * k);
    }
    else if(op->type == ORC_CHILD_PEEL) {
        transmission_matrix(c, op->peelnode, c->tmat);
    }

    for(long idx = 0; idx < size; ++idx) {
        int valid = 1;
        for(int d = 0; d < op->ncut; ++d) {
            int v = (int)((idx >> (2 * d)) & 3);
            c->assign[op->cutset[d]] = v;
            if(!orc_is_legal(p, op->cutset[d], c->locus, v)) { valid = 0; break; }
        }
        if(valid) sampler_element(c, opi, idx, trait_cache);
    }
}

static double sampler_forward(orc_ctx* c) {
    for(int i = 0; i < c->p->nops; ++i) sampler_evaluate(c, i);
    return c->mat[c->p->nops - 1][0];                             /* peel_matrix.cc:70-77 */
}

double orc_ls_forward(const orc_problem* p, const int* dg, int locus, int ignore_left, int ignore_right,
                      double* matrices, double* presums) {
    orc_ctx* c = ctx_new(p, 1);
    double res;
    long mo = 0, po = 0;
    c->dg = dg;
    sampler_set_locus(c, locus, ignore_left, ignore_right);
    res = sampler_forward(c);
    for(int i = 0; i < p->nops; ++i) {
        long n = ipow4(p->ops[i].ncut);
        if(matrices) memcpy(matrices + mo, c->mat[i], sizeof(double) * n);
        if(presums)  memcpy(presums + po, c->pre[i], sizeof(double) * n * 4);
        mo += n; po += 4 * n;
    }
    ctx_free(c);
    return res;
}

/* locus_sampler2.cc:44-65 */
static double homo_p0(const orc_ctx* c, int person, int parent) {
    const orc_problem* p = c->p;
    double d0 = 1.0, d1 = 1.0;
    if(c->locus != 0 && !c->ignore_left) {
        int v = dg_get(p, c->dg, person, c->locus - 1, parent);
        double th = p->theta[c->locus - 1], ith = 1.0 - p->theta[c->locus - 1];
        d0 *= (v == 0) ? ith : th;
        d1 *= (v == 1) ? ith : th;
    }
    if(c->locus != p->M - 1 && !c->ignore_right) {
        int v = dg_get(p, c->dg, person, c->locus + 1, parent);
        double th = p->theta[c->locus], ith = 1.0 - p->theta[c->locus];
        d0 *= (v == 0) ? ith : th;
        d1 *= (v == 1) ? ith : th;
    }
    return d0 / (d0 + d1);
}

double orc_homo_p0(const orc_problem* p, const int* dg, int locus, int person, int parent,
                   int ignore_left, int ignore_right) {
    orc_ctx c;
    memset(&c, 0, sizeof(c));
    c.p = p; c.dg = dg; c.locus = locus; c.ignore_left = ignore_left; c.ignore_right = ignore_right;
    return homo_p0(&c, person, parent);
}

/* locus_sampler2.cc:32-39 */
static int hetero_mi(int allele, int trait) {
    if(allele == 0) return (trait == ORC_UA) ? 0 : 1;
    return (trait == ORC_UA) ? 1 : 0;
}

double orc_ls_step(const orc_problem* p, int* dg, int locus, int ignore_left, int ignore_right,
                   uint64_t seed, uint32_t chain, uint64_t iteration, int* pmk_out, double* dist4_out) {
    orc_ctx* c = ctx_new(p, 1);
    double res;
    int* pmk = (int*) malloc(sizeof(int) * p->N);
    c->dg = dg;
    sampler_set_locus(c, locus, ignore_left, ignore_right);
    res = sampler_forward(c);
    if(res == 0.0) {                                              /* locus_sampler2.cc:137-142 */
        free(pmk); ctx_free(c);
        return 0.0;
    }

    for(int i = 0; i < p->N; ++i) pmk[i] = -1;

    /* backward: sampler_rfunction.cc:159-188 and rfunction.cc:200-209 */
    for(int i = p->nops - 1; i >= 0; --i) {
        const orc_op* op = &p->ops[i];
        double d[4], total, cum = 0.0, r;
        long idx = 0;
        int last = 0, chosen = -1;
        for(int k = 0; k < op->ncut; ++k) idx += (long) pmk[op->cutset[k]] << (2 * k);
        for(int g = 0; g < 4; ++g) d[g] = c->pre[i][idx + ipow4(op->ncut) * g];
        if(dist4_out) memcpy(dist4_out + 4 * i, d, sizeof(d));
        total = d[0] + d[1] + d[2] + d[3];
        if(total != 0.0) for(int g = 0; g < 4; ++g) d[g] /= total;
        r = orc_uniform(seed, chain, iteration, (uint32_t)(locus + g_locus_key_offset), (uint32_t) i);
        for(int g = 0; g < 4; ++g) {
            cum += d[g];
            if(r < cum) { chosen = g; break; }
            if(d[g] != 0.0) last = g;
        }
        pmk[op->peelnode] = (chosen >= 0) ? chosen : last;
    }

    /* meiosis indicators: locus_sampler2.cc:93-125 */
    for(int i = 0; i < p->N; ++i) {
        int trait, mt, ma, mi;
        if(is_founder(p, i)) continue;
        trait = pmk[i];
        mt = pmk[p->mother[i]];
        ma = (trait == ORC_UU || trait == ORC_UA) ? 0 : 1;
        if(mt == ORC_UA || mt == ORC_AU) mi = hetero_mi(ma, mt);
        else {
            double p0 = homo_p0(c, i, MATERNAL);
            mi = (orc_uniform(seed, chain, iteration, (uint32_t)(locus + g_locus_key_offset), (uint32_t)(p->nops + 2 * i + MATERNAL)) < p0) ? 0 : 1;
        }
        dg[((long) locus * p->N + i) * 2 + MATERNAL] = mi;

        if(p->sex_linked) {
            dg[((long) locus * p->N + i) * 2 + PATERNAL] = MATERNAL;    /* :115-118 */
            continue;
        }
        {
            int pt = pmk[p->father[i]];
            int pa = (trait == ORC_UU || trait == ORC_AU) ? 0 : 1;
            int pi;
            if(pt == ORC_UA || pt == ORC_AU) pi = hetero_mi(pa, pt);
            else {
                double p0 = homo_p0(c, i, PATERNAL);
                pi = (orc_uniform(seed, chain, iteration, (uint32_t)(locus + g_locus_key_offset), (uint32_t)(p->nops + 2 * i + PATERNAL)) < p0) ? 0 : 1;
            }
            dg[((long) locus * p->N + i) * 2 + PATERNAL] = pi;
        }
    }

    if(pmk_out) memcpy(pmk_out, pmk, sizeof(int) * p->N);
    free(pmk);
    ctx_free(c);
    return res;
}

int orc_ls_sweep(const orc_problem* p, int* dg, uint64_t seed, uint32_t chain, uint64_t iteration) {
    int first = (orc_uniform(seed, chain, iteration, 0u, ORC_SLOT_PHASE) < 0.5) ? 0 : 1;
    for(int phase = 0; phase < 2; ++phase) {
        int start = (phase == 0) ? first : 1 - first;
        for(int l = start; l < p->M; l += 2) {
            double res = orc_ls_step(p, dg, l, 0, 0, seed, chain, iteration, 0, 0);
            if(res == 0.0) return 1 + l;
        }
    }
    return 0;
}

double orc_si_start_from(const orc_problem* p, int* dg, int start_locus, uint64_t seed, uint32_t chain, uint64_t run) {
    double weight = 0.0, res;
    res = orc_ls_step(p, dg, start_locus, 1, 1, seed, chain, run, 0, 0);
    if(res == 0.0) return LOG_ZERO;
    weight += log(res);
    for(int l = start_locus - 1; l >= 0; --l) {
        res = orc_ls_step(p, dg, l, 1, 0, seed, chain, run, 0, 0);
        if(res == 0.0) return LOG_ZERO;
        weight += log(res);
    }
    for(int l = start_locus + 1; l < p->M; ++l) {
        res = orc_ls_step(p, dg, l, 0, 1, seed, chain, run, 0, 0);
        if(res == 0.0) return LOG_ZERO;
        weight += log(res);
    }
    return weight;
}

/* One ELOD replicate (elod.cc:55-61): LocusSampler::start_from(dg1, 1) on the three-locus problem,
   DescentGraph::copy_locus of loci 0 and 2 into a two-locus graph, Peeler::process on the two-locus
   problem.  dg3 int[3][N][2] receives the sampled graph; returns ln-prob (LOG_ZERO if the peel is <= 0). */
double orc_elod_replicate(const orc_problem* p1, const orc_problem* p2, long replicate, uint64_t seed, uint32_t chain, int* dg3) {
    const int N = p1->N;
    int* dg2 = (int*) calloc((size_t) 2 * N * 2, sizeof(int));
    double result[1], prob[1], w;
    memset(dg3, 0, sizeof(int) * 3 * N * 2);
    g_locus_key_offset = 3 * replicate;
    w = orc_si_start_from(p1, dg3, 1, seed, chain, 0);
    g_locus_key_offset = 0;
    (void) w;
    memcpy(dg2, dg3, sizeof(int) * 2 * N);
    memcpy(dg2 + 2 * N, dg3 + 4 * N, sizeof(int) * 2 * N);
    orc_lod_interval(p2, dg2, 0, result, prob, -1, 0);
    free(dg2);
    return result[0] > 0.0 ? prob[0] : LOG_ZERO;
}

/* ---- trait R-function -------------------------------------------------------------- */

/* trait_rfunction.cc:9-22 */
static double trait_recombination(const orc_ctx* c, int person, int mat_allele, int pat_allele) {
    const orc_problem* p = c->p;
    double tmp = 1.0;
    tmp *= (dg_get(p, c->dg, person, c->locus,     MATERNAL) == mat_allele) ? c->antitheta  : c->theta;
    tmp *= (dg_get(p, c->dg, person, c->locus + 1, MATERNAL) == mat_allele) ? c->antitheta2 : c->theta2;
    if(!p->sex_linked) {
        tmp *= (dg_get(p, c->dg, person, c->locus,     PATERNAL) == pat_allele) ? c->antitheta  : c->theta;
        tmp *= (dg_get(p, c->dg, person, c->locus + 1, PATERNAL) == pat_allele) ? c->antitheta2 : c->theta2;
    }
    return tmp;
}

/* trait_rfunction.cc:28-69, :71-137 and rfunction.cc:116-142 */
static void trait_element(orc_ctx* c, int opi, long idx, const double trait_cache[4]) {
    const orc_problem* p = c->p;
    const orc_op* op = &p->ops[opi];
    double trait_prob = p->sex_linked ? 0.5 : 0.25;
    double total = 0.0;

    if(op->type == ORC_CHILD_PEEL) {
        int kid = op->peelnode;
        int m = c->assign[p->mother[kid]];
        int f = c->assign[p->father[kid]];
        for(int i = 0; i < 2; ++i) {
            for(int j = 0; j < 2; ++j) {
                int kt = phased_trait(p, m, f, i, j, p->sex[kid]);
                double tmp = trait_cache[kt];
                c->assign[kid] = kt;
                if(tmp == 0.0) continue;
                for(int k = 0; k < op->nprev; ++k) tmp *= prev_get(c, op->prev[k]);
                tmp *= (!c->dg) ? trait_prob : trait_prob * trait_recombination(c, kid, i, j);
                total += tmp;
            }
        }
    }
    else if(op->type == ORC_PARENT_PEEL) {
        for(int a = 0; a < 4; ++a) {
            double tmp = trait_cache[a];
            double child_prob = 1.0;
            int m = a, f = a;
            c->assign[op->peelnode] = a;
            if(tmp == 0.0) continue;
            for(int k = 0; k < op->nprev; ++k) tmp *= prev_get(c, op->prev[k]);
            for(int d = 0; d < op->ncut; ++d) {
                int kid = op->cutset[d];
                int kid_trait = c->assign[kid];
                double child_tmp = 0.0;
                if(!(p->mother[kid] == op->peelnode || p->father[kid] == op->peelnode)) continue;
                if(p->mother[kid] == op->peelnode) f = c->assign[p->father[kid]];
                else                               m = c->assign[p->mother[kid]];
                for(int i = 0; i < 2; ++i) {
                    for(int j = 0; j < 2; ++j) {
                        if(phased_trait(p, m, f, i, j, p->sex[kid]) != kid_trait) continue;
                        child_tmp += (!c->dg) ? trait_prob : trait_prob * trait_recombination(c, kid, i, j);
                    }
                }
                child_prob *= child_tmp;
            }
            tmp *= child_prob;
            total += tmp;
        }
    }
    else {
        for(int g = 0; g < 4; ++g) {
            double tmp = trait_cache[g];
            c->assign[op->peelnode] = g;
            if(tmp == 0.0) continue;
            for(int k = 0; k < op->nprev; ++k) tmp *= prev_get(c, op->prev[k]);
            total += tmp;
        }
    }
    c->mat[opi][idx] = total;
}

/* rfunction.cc:172-198 (offset != 0 branch): only valid_lod_indices are visited; the matrices
   are never reset, every other cell keeps its construction-time 0 */
static void trait_evaluate(orc_ctx* c, int opi) {
    const orc_problem* p = c->p;
    const orc_op* op = &p->ops[opi];
    long size = ipow4(op->ncut);
    const double* trait_cache = p->disease_prob + 4 * op->peelnode;
    for(long idx = 0; idx < size; ++idx) {
        int valid = 1;
        for(int d = 0; d < op->ncut; ++d) {
            int v = (int)((idx >> (2 * d)) & 3);
            c->assign[op->cutset[d]] = v;
            if(p->disease_prob[op->cutset[d] * 4 + v] == 0.0) { valid = 0; break; }
        }
        if(valid) trait_element(c, opi, idx, trait_cache);
    }
}

/* trait_rfunction.h:50-56 */
static void trait_set_thetas(orc_ctx* c, int interval, int offset) {
    const orc_problem* p = c->p;
    c->locus = interval;
    c->theta = p->partial[interval] * offset;                     /* genetic_map.cc:124-126 */
    c->antitheta = 1.0 - c->theta;
    c->theta2 = p->partial[interval] * (p->nlod + 1 - offset);
    c->antitheta2 = 1.0 - c->theta2;
}

/* descent_graph.cc:212-242 */
double orc_recombination_prob(const orc_problem* p, const int* dg, int locus) {
    double tmp = 0.0;
    double lt = log(p->theta[locus]);                             /* genetic_map.cc:132-138 */
    double lit = log(1.0 - p->theta[locus]);
    int nall = p->sex_linked ? 1 : 2;
    for(int i = 0; i < p->N; ++i) {
        if(is_founder(p, i)) continue;
        for(int j = 0; j < nall; ++j) {
            int cross = dg_get(p, dg, i, locus, j) != dg_get(p, dg, i, locus + 1, j);
            tmp += cross ? lt : lit;
        }
    }
    return tmp;
}

/* descent_graph.cc:22,35 */
double orc_marker_transmission(const orc_problem* p) {
    int nf = p->N - p->F;
    return p->sex_linked ? log(0.5) * nf : log(0.5) * (2 * nf);
}

void orc_lod_interval(const orc_problem* p, const int* dg, int interval, double* result, double* prob,
                      int dump_k, double* matrices) {
    orc_ctx* c = ctx_new(p, 0);
    c->dg = dg;
    c->trait_mode = 1;
    for(int k = 0; k < p->nlod; ++k) {
        double res;
        trait_set_thetas(c, interval, k + 1);
        for(int i = 0; i < p->nops; ++i) trait_evaluate(c, i);
        res = c->mat[p->nops - 1][0];
        result[k] = res;
        prob[k] = log(res) - orc_recombination_prob(p, dg, interval) - orc_marker_transmission(p);
        if(matrices && k == dump_k) {
            long mo = 0;
            for(int i = 0; i < p->nops; ++i) {
                long n = ipow4(p->ops[i].ncut);
                memcpy(matrices + mo, c->mat[i], sizeof(double) * n);
                mo += n;
            }
        }
    }
    ctx_free(c);
}

double orc_trait_prob(const orc_problem* p) {
    orc_ctx* c = ctx_new(p, 0);
    double res;
    c->dg = 0;
    c->trait_mode = 1;
    c->locus = 0;
    for(int i = 0; i < p->nops; ++i) trait_evaluate(c, i);
    res = log(c->mat[p->nops - 1][0]);
    ctx_free(c);
    return res;
}

/* ---- LOD accumulators (logarithms.cc:14-29, lod_score.h:74-88) --------------------- */

double orc_log_sum(double a, double b) {
    if(a == LOG_ZERO) return b;
    if(b == LOG_ZERO) return a;
    return log(exp(b - a) + 1) + a;
}

void orc_lod_add(double* scores, int n, const double* prob, int first) {
    for(int i = 0; i < n; ++i) scores[i] = first ? prob[i] : orc_log_sum(prob[i], scores[i]);
}

double orc_lod_normalise(double score, int count, double trait_prob) {
    return (score - log((double) count) - trait_prob) / log(10.0);
}

void orc_lod_pass(const orc_problem* p, const int* dg, double* scores, int first) {
    double* res = (double*) malloc(sizeof(double) * p->nlod);
    double* prob = (double*) malloc(sizeof(double) * p->nlod);
    for(int l = 0; l < p->M - 1; ++l) {
        orc_lod_interval(p, dg, l, res, prob, -1, 0);
        orc_lod_add(scores + (long) l * p->nlod, p->nlod, prob, first);
    }
    free(res); free(prob);
}

/* ---- RNG exports ------------------------------------------------------------------- */

void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    orc_philox4x32_10(ctr, key, out);
}

double orc_uniform_draw(uint64_t seed, uint32_t chain, uint64_t iteration, uint32_t locus, uint32_t slot) {
    return orc_uniform(seed, chain, iteration, locus, slot);
}
