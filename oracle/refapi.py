"""ctypes binding to oracle/_ref/libswiftref.so -- TEST INFRASTRUCTURE ONLY.

The library is the UNMODIFIED reference (ajm/swiftlink) compiled by oracle/Makefile plus the
dump driver oracle/ref_driver.cc.  Only tests/, tests/golden/make_golden.py, bench.py's
reference/cpu_baseline legs and __graft_entry__.smoke() may import this module; the product
package (swiftlink_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
LIB_PATH = os.path.join(REF_DIR, "libswiftref.so")
EXAMPLES = os.path.join(REF_DIR, "examples")

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_open.restype = C.c_void_p
        L.ref_open.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        L.ref_close.argtypes = [C.c_void_p]
        L.ref_peel_cost.restype = C.c_uint
        for name in ("ref_dg_likelihood", "ref_dg_recombination_prob", "ref_dg_marker_transmission",
                     "ref_chain_run", "ref_ls_forward", "ref_calc_trait_prob", "ref_bench_lsweeps",
                     "ref_bench_lodpasses", "ref_fag", "ref_fag_flipped", "ref_bench_msweeps", "ref_elod", "ref_elod2", "ref_bench_chain"):
            getattr(L, name).restype = C.c_double
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _ip(a):
    return _p(a, C.c_int)


def _dp(a):
    return _p(a, C.c_double)


def set_threads(n):
    lib().ref_set_threads(C.c_int(n))


def seed(s):
    lib().ref_seed(C.c_uint(s))


def elod(pedfile, frequency=1e-4, penetrance=(0.0, 0.0, 1.0), separation=0.05, replicates=10000, sex_linked=False, seed=1,
         affected_only=False):
    """the reference's Elod(pedfile, options).run()"""
    pen = np.ascontiguousarray(penetrance, np.float64)
    return float(lib().ref_elod2(pedfile.encode(), C.c_double(frequency), _dp(pen), C.c_double(separation), int(replicates),
                                 int(sex_linked), int(affected_only), C.c_uint(seed)))


def example(name):
    return tuple(os.path.join(EXAMPLES, "%s.%s" % (name, ext)) for ext in ("ped", "map", "dat"))


class Ref(object):
    """One pedigree loaded through the reference's own parsers."""

    def __init__(self, ped, mapf, dat, sex_linked=False, lodscores=5):
        self.L = lib()
        self.h = self.L.ref_open(ped.encode(), mapf.encode(), dat.encode(), int(sex_linked), int(lodscores))
        if not self.h:
            raise RuntimeError("reference failed to parse %s" % ped)
        self.h = C.c_void_p(self.h)
        d = np.zeros(6, dtype=np.int32)
        self.L.ref_dims(self.h, _ip(d))
        self.N, self.F, self.M, self.nlod, self.sex_linked, self.leaves = [int(x) for x in d]

    def close(self):
        if self.h:
            self.L.ref_close(self.h)
            self.h = None

    # ---- inputs -------------------------------------------------------------------
    def person_table(self):
        N = self.N
        mother = np.zeros(N, np.int32); father = np.zeros(N, np.int32); sex = np.zeros(N, np.int32)
        aff = np.zeros(N, np.int32); typed = np.zeros(N, np.int32); dprob = np.zeros((N, 4))
        self.L.ref_person_table(self.h, _ip(mother), _ip(father), _ip(sex), _ip(aff), _ip(typed), _dp(dprob))
        return dict(mother=mother, father=father, sex=sex, affection=aff, typed=typed, disease_prob=dprob)

    def person_names(self):
        out = []
        buf = C.create_string_buffer(256)
        for i in range(self.N):
            self.L.ref_person_name(self.h, i, buf, 256)
            out.append(buf.value.decode())
        return out

    def marker_names(self):
        out = []
        buf = C.create_string_buffer(256)
        for i in range(self.M):
            self.L.ref_marker_name(self.h, i, buf, 256)
            out.append(buf.value.decode())
        return out

    def genotypes(self):
        g = np.zeros((self.N, self.M), np.int32)
        self.L.ref_genotypes(self.h, _ip(g))
        return g

    def marker_trait_prob(self):
        t = np.zeros((self.N, self.M, 4))
        self.L.ref_marker_trait_prob(self.h, _dp(t))
        return t

    def map_table(self):
        M = self.M
        gdist = np.zeros(M); minor = np.zeros(M); prob = np.zeros((M, 4)); xprob = np.zeros((M, 4))
        theta = np.zeros(M - 1); partial = np.zeros(M - 1)
        self.L.ref_map_table(self.h, _dp(gdist), _dp(minor), _dp(prob), _dp(xprob), _dp(theta), _dp(partial))
        return dict(gdist=gdist, minor=minor, prob=prob, xprob=xprob, theta=theta, partial=partial)

    def disease_model(self):
        d = np.zeros(4)
        self.L.ref_disease_model(self.h, _dp(d))
        return dict(freq=float(d[0]), penetrance=d[1:].copy())

    # ---- plan ---------------------------------------------------------------------
    def build_peel(self, iterations=1000000):
        self.L.ref_build_peel(self.h, int(iterations))

    def set_peel(self, seq):
        s = np.ascontiguousarray(seq, dtype=np.uint32)
        assert s.shape == (self.N,)
        return bool(self.L.ref_set_peel(self.h, _p(s, C.c_uint)))

    def elim_masks(self):
        m = np.zeros((self.M, self.N), np.int32)
        self.L.ref_elim_masks(self.h, _ip(m))
        return m

    def num_ops(self):
        return int(self.L.ref_num_ops(self.h))

    def peel_cost(self):
        return int(self.L.ref_peel_cost(self.h))

    def ops(self):
        out = []
        N = self.N
        for i in range(self.num_ops()):
            info = np.zeros(5, np.int32); cut = np.zeros(N, np.int32); prev = np.zeros(N, np.int32)
            kids = np.zeros(N, np.int32)
            self.L.ref_op_info(self.h, i, _ip(info), _ip(cut), _ip(prev), _ip(kids))
            out.append(dict(type=int(info[0]), peelnode=int(info[1]), cutset=cut[:info[2]].tolist(),
                            previous=prev[:info[3]].tolist(), children=kids[:info[4]].tolist()))
        return out

    def op_indices(self, i, which, locus=0):
        n = int(self.L.ref_op_indices(self.h, i, which, locus, None, 0))
        buf = np.zeros(max(n, 1), np.int32)
        self.L.ref_op_indices(self.h, i, which, locus, _ip(buf), n)
        return buf[:n].copy()

    # ---- descent graph --------------------------------------------------------------
    def dg_random(self):
        return bool(self.L.ref_dg_random(self.h))

    def dg_get(self):
        g = np.zeros((self.M, self.N, 2), np.int32)
        self.L.ref_dg_get(self.h, _ip(g))
        return g

    def dg_set(self, g):
        g = np.ascontiguousarray(g, dtype=np.int32)
        assert g.shape == (self.M, self.N, 2)
        self.L.ref_dg_set(self.h, _ip(g))

    def dg_likelihood(self):
        return float(self.L.ref_dg_likelihood(self.h))

    def dg_recombination_prob(self, locus):
        return float(self.L.ref_dg_recombination_prob(self.h, int(locus)))

    def dg_marker_transmission(self):
        return float(self.L.ref_dg_marker_transmission(self.h))

    def sequential_imputation(self, iterations):
        self.L.ref_sequential_imputation(self.h, int(iterations))

    def chain_run(self, burnin, iterations, scoring_period=10, lsampler_prob=0.5):
        n = (self.M - 1) * self.nlod
        raw = np.zeros(n); lod = np.zeros(n); cnt = C.c_int(0)
        tp = self.L.ref_chain_run(self.h, int(burnin), int(iterations), int(scoring_period),
                                  C.c_double(lsampler_prob), _dp(raw), _dp(lod), C.byref(cnt))
        return dict(trait_prob=float(tp), raw=raw.reshape(self.M - 1, self.nlod),
                    lod=lod.reshape(self.M - 1, self.nlod), count=cnt.value)

    # ---- L-sampler ------------------------------------------------------------------
    def matrix_sizes(self):
        cs = [len(o["cutset"]) for o in self.ops()]
        return [4 ** c for c in cs], [4 ** (c + 1) for c in cs]

    def ls_forward(self, locus, mode=0, ignore_left=False, ignore_right=False):
        ms, ps = self.matrix_sizes()
        mat = np.zeros(sum(ms)); pre = np.zeros(sum(ps))
        res = self.L.ref_ls_forward(self.h, int(locus), int(mode), int(ignore_left), int(ignore_right),
                                    _dp(mat), _dp(pre))
        return float(res), mat, pre

    def ls_step(self, locus):
        self.L.ref_ls_step(self.h, int(locus))

    def ls_backward_trace(self):
        pmk = np.zeros(self.N, np.int32); dist = np.zeros((self.num_ops(), 4))
        self.L.ref_ls_backward_trace(self.h, _ip(pmk), _dp(dist))
        return pmk, dist

    def ls_sample_indicators(self, pmk):
        p = np.ascontiguousarray(pmk, dtype=np.int32)
        self.L.ref_ls_sample_indicators(self.h, _ip(p))

    # ---- M-sampler --------------------------------------------------------------------
    def fag(self, locus, flip=None):
        """(edge_list[2N], likelihood) of the founder allele graph at `locus`; flip = (person, parent)
        applies FounderAlleleGraph4::flip first"""
        edge = np.zeros(2 * self.N, np.int32)
        if flip is None:
            lik = self.L.ref_fag(self.h, int(locus), _ip(edge))
        else:
            lik = self.L.ref_fag_flipped(self.h, int(locus), int(flip[0]), int(flip[1]), _ip(edge))
        return edge, float(lik)

    def ms_ordering(self):
        out = np.zeros(2 * self.N, np.int32)
        n = int(self.L.ref_ms_ordering(self.h, _ip(out)))
        return out[:n].copy()

    def ms_reset(self, parameter):
        self.L.ref_ms_reset(self.h, int(parameter))

    def ms_raw(self):
        raw = np.zeros((self.M, 2))
        self.L.ref_ms_raw(self.h, _dp(raw))
        return raw

    def ms_step(self, parameter):
        """returns (uniforms the step had available, raw_matrix, fb_matrix) -- the context's descent
        graph is updated in place"""
        us = np.zeros(self.M); raw = np.zeros((self.M, 2)); fb = np.zeros((self.M, 2))
        self.L.ref_ms_step(self.h, int(parameter), _dp(us), self.M, _dp(raw), _dp(fb))
        return us, raw, fb

    def bench_msweeps(self, reps):
        return float(self.L.ref_bench_msweeps(self.h, int(reps)))

    # ---- LOD ------------------------------------------------------------------------
    def calc_trait_prob(self):
        return float(self.L.ref_calc_trait_prob(self.h))

    def lod_interval(self, interval, dump_k=-1):
        res = np.zeros(self.nlod); prob = np.zeros(self.nlod)
        mat = None
        if dump_k >= 0:
            ms, _ = self.matrix_sizes()
            mat = np.zeros(sum(ms))
        self.L.ref_lod_interval(self.h, int(interval), _dp(res), _dp(prob), int(dump_k),
                                _dp(mat) if mat is not None else None)
        return res, prob, mat

    # ---- timing ---------------------------------------------------------------------
    def bench_lsweeps(self, reps, lgroups=-1):
        return float(self.L.ref_bench_lsweeps(self.h, int(reps), int(lgroups)))

    def bench_lodpasses(self, reps):
        return float(self.L.ref_bench_lodpasses(self.h, int(reps)))

    def bench_chain(self, iterations, scoring_period, lsampler_prob, lod_intervals=0, sched_trials=1):
        """MarkovChain::run's loop over the reference's own objects; returns (seconds, lgroups, [L-sweeps, M-sweeps, scorings])"""
        lg = C.c_int(0)
        counts = np.zeros(3, np.int32)
        t = float(self.L.ref_bench_chain(self.h, int(iterations), int(scoring_period), C.c_double(lsampler_prob),
                                         int(lod_intervals), int(sched_trials), C.byref(lg), _p(counts, C.c_int32)))
        return t, int(lg.value), [int(x) for x in counts]
