"""ctypes binding to oracle/liboracle.so (oracle/peel_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, tests/golden/make_golden.py, bench.py's cpu_baseline/reference legs and
__graft_entry__.smoke() may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
MAXC = 12


class OrcOp(C.Structure):
    _fields_ = [("type", C.c_int), ("peelnode", C.c_int),
                ("ncut", C.c_int), ("cutset", C.c_int * MAXC),
                ("nprev", C.c_int), ("prev", C.c_int * MAXC),
                ("nchild", C.c_int), ("children", C.c_int * MAXC)]


class OrcProblem(C.Structure):
    _fields_ = [("N", C.c_int), ("F", C.c_int), ("M", C.c_int), ("nlod", C.c_int), ("sex_linked", C.c_int),
                ("mother", C.POINTER(C.c_int)), ("father", C.POINTER(C.c_int)), ("sex", C.POINTER(C.c_int)),
                ("disease_prob", C.POINTER(C.c_double)), ("marker_prob", C.POINTER(C.c_double)),
                ("elim", C.POINTER(C.c_int)), ("theta", C.POINTER(C.c_double)),
                ("partial", C.POINTER(C.c_double)),
                ("nops", C.c_int), ("ops", C.POINTER(OrcOp)),
                ("typed", C.POINTER(C.c_int)), ("genotypes", C.POINTER(C.c_int)), ("minor", C.POINTER(C.c_double))]


_lib = None


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < max(os.path.getmtime(os.path.join(HERE, f))
                                             for f in ("peel_oracle.c", "msampler_oracle.c", "peel_oracle.h", "philox.h")):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        for name in ("orc_ls_forward", "orc_ls_step", "orc_homo_p0", "orc_si_start_from", "orc_trait_prob",
                     "orc_recombination_prob", "orc_marker_transmission", "orc_log_sum",
                     "orc_lod_normalise", "orc_uniform_draw", "orc_fag_likelihood", "orc_dg_sum_prior_prob",
                     "orc_dg_likelihood", "orc_elod_replicate"):
            getattr(L, name).restype = C.c_double
        L.orc_matrix_doubles.restype = C.c_long
        L.orc_presum_doubles.restype = C.c_long
        L.orc_ms_create.restype = C.c_void_p
        L.orc_ms_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Problem(object):
    """Flat-array problem description (the reference's own shapes).

    Required keys of `d`: N, F, M, nlod, sex_linked, mother, father, sex, disease_prob [N,4],
    marker_prob [N,M,4], elim [M,N], theta [M-1], partial [M-1], ops (list of dicts with
    type, peelnode, cutset, previous, children).
    """

    def __init__(self, d):
        self.L = lib()
        self.N, self.F, self.M, self.nlod = int(d["N"]), int(d["F"]), int(d["M"]), int(d["nlod"])
        self.sex_linked = int(d["sex_linked"])
        self._keep = {}
        for k, dt in (("mother", np.int32), ("father", np.int32), ("sex", np.int32), ("elim", np.int32),
                      ("disease_prob", np.float64), ("marker_prob", np.float64), ("theta", np.float64),
                      ("partial", np.float64), ("typed", np.int32), ("genotypes", np.int32), ("minor", np.float64)):
            self._keep[k] = np.ascontiguousarray(d[k], dtype=dt)
        ops = d["ops"]
        self.nops = len(ops)
        arr = (OrcOp * self.nops)()
        for i, o in enumerate(ops):
            arr[i].type = int(o["type"])
            arr[i].peelnode = int(o["peelnode"])
            for name, cnt, key in (("cutset", "ncut", "cutset"), ("prev", "nprev", "previous"),
                                   ("children", "nchild", "children")):
                vals = [int(x) for x in o[key]]
                assert len(vals) <= MAXC
                setattr(arr[i], cnt, len(vals))
                for j, v in enumerate(vals):
                    getattr(arr[i], name)[j] = v
        self._ops = arr
        self.ops = ops
        k = self._keep
        self.c = OrcProblem(self.N, self.F, self.M, self.nlod, self.sex_linked,
                            _ip(k["mother"]), _ip(k["father"]), _ip(k["sex"]),
                            _dp(k["disease_prob"]), _dp(k["marker_prob"]), _ip(k["elim"]),
                            _dp(k["theta"]), _dp(k["partial"]), self.nops, arr,
                            _ip(k["typed"]), _ip(k["genotypes"]), _dp(k["minor"]))
        self.p = C.byref(self.c)
        self.matrix_doubles = int(self.L.orc_matrix_doubles(self.p))
        self.presum_doubles = int(self.L.orc_presum_doubles(self.p))

    # ---- index tables -----------------------------------------------------------------
    def op_indices(self, op, which, locus=0):
        n = int(self.L.orc_op_indices(self.p, int(op), int(which), int(locus), None, 0))
        buf = np.zeros(max(n, 1), np.int32)
        self.L.orc_op_indices(self.p, int(op), int(which), int(locus), _ip(buf), n)
        return buf[:n].copy()

    # ---- L-sampler ----------------------------------------------------------------------
    def ls_forward(self, dg, locus, ignore_left=False, ignore_right=False):
        dg = np.ascontiguousarray(dg, dtype=np.int32)
        mat = np.zeros(self.matrix_doubles); pre = np.zeros(self.presum_doubles)
        res = self.L.orc_ls_forward(self.p, _ip(dg), int(locus), int(ignore_left), int(ignore_right),
                                    _dp(mat), _dp(pre))
        return float(res), mat, pre

    def ls_step(self, dg, locus, seed, chain, iteration, ignore_left=False, ignore_right=False):
        """dg (int32 [M,N,2]) is updated in place; returns (likelihood, pmk, dist4)."""
        assert dg.dtype == np.int32 and dg.flags["C_CONTIGUOUS"]
        pmk = np.zeros(self.N, np.int32); dist = np.zeros((self.nops, 4))
        res = self.L.orc_ls_step(self.p, _ip(dg), int(locus), int(ignore_left), int(ignore_right),
                                 C.c_uint64(seed), C.c_uint32(chain), C.c_uint64(iteration),
                                 _ip(pmk), _dp(dist))
        return float(res), pmk, dist

    def ls_sweep(self, dg, seed, chain, iteration):
        assert dg.dtype == np.int32 and dg.flags["C_CONTIGUOUS"]
        return int(self.L.orc_ls_sweep(self.p, _ip(dg), C.c_uint64(seed), C.c_uint32(chain),
                                       C.c_uint64(iteration)))

    def si_start_from(self, dg, start_locus, seed, chain, run):
        assert dg.dtype == np.int32 and dg.flags["C_CONTIGUOUS"]
        return float(self.L.orc_si_start_from(self.p, _ip(dg), int(start_locus), C.c_uint64(seed),
                                              C.c_uint32(chain), C.c_uint64(run)))

    def homo_p0(self, dg, locus, person, parent, ignore_left=False, ignore_right=False):
        dg = np.ascontiguousarray(dg, dtype=np.int32)
        return float(self.L.orc_homo_p0(self.p, _ip(dg), int(locus), int(person), int(parent),
                                        int(ignore_left), int(ignore_right)))

    # ---- M-sampler ----------------------------------------------------------------------
    def fag(self, dg, locus, flip=None):
        dg = np.ascontiguousarray(dg, dtype=np.int32)
        edge = np.zeros(2 * self.N, np.int32)
        self.L.orc_fag_reset(self.p, _ip(dg), int(locus), _ip(edge))
        if flip is not None:
            self.L.orc_fag_flip(self.p, _ip(dg), int(locus), int(flip[0]), int(flip[1]), _ip(edge))
        return edge, float(self.L.orc_fag_likelihood(self.p, int(locus), _ip(edge)))

    def ms_ordering(self):
        out = np.zeros(2 * self.N + 1, np.int32)
        n = int(self.L.orc_ms_ordering(self.p, _ip(out)))
        return out[:n].copy()

    def ms_shuffle(self, order, seed, chain, iteration):
        v = np.ascontiguousarray(order, dtype=np.int32).copy()
        self.L.orc_ms_shuffle(_ip(v), len(v), C.c_uint64(seed), C.c_uint32(chain), C.c_uint64(iteration))
        return v

    def msampler(self):
        return MSampler(self)

    def ms_sweep(self, dg, seed, chain, iteration):
        assert dg.dtype == np.int32 and dg.flags["C_CONTIGUOUS"]
        return int(self.L.orc_ms_sweep(self.p, _ip(dg), C.c_uint64(seed), C.c_uint32(chain), C.c_uint64(iteration)))

    def dg_likelihood(self, dg):
        dg = np.ascontiguousarray(dg, dtype=np.int32)
        return float(self.L.orc_dg_likelihood(self.p, _ip(dg)))

    def dg_sum_prior_prob(self, dg):
        dg = np.ascontiguousarray(dg, dtype=np.int32)
        return float(self.L.orc_dg_sum_prior_prob(self.p, _ip(dg)))

    # ---- LOD ----------------------------------------------------------------------------
    def lod_interval(self, dg, interval, dump_k=-1):
        dg = np.ascontiguousarray(dg, dtype=np.int32)
        res = np.zeros(self.nlod); prob = np.zeros(self.nlod)
        mat = np.zeros(self.matrix_doubles) if dump_k >= 0 else None
        self.L.orc_lod_interval(self.p, _ip(dg), int(interval), _dp(res), _dp(prob), int(dump_k),
                                _dp(mat) if mat is not None else None)
        return res, prob, mat

    def trait_prob(self):
        return float(self.L.orc_trait_prob(self.p))

    def recombination_prob(self, dg, locus):
        dg = np.ascontiguousarray(dg, dtype=np.int32)
        return float(self.L.orc_recombination_prob(self.p, _ip(dg), int(locus)))

    def marker_transmission(self):
        return float(self.L.orc_marker_transmission(self.p))

    def lod_pass(self, dg, scores, first):
        dg = np.ascontiguousarray(dg, dtype=np.int32)
        assert scores.dtype == np.float64 and scores.size == (self.M - 1) * self.nlod
        self.L.orc_lod_pass(self.p, _ip(dg), _dp(scores), int(first))


class MSampler(object):
    """MeiosisSampler restatement (oracle/msampler_oracle.c); dg arrays are int32 [M,N,2], updated in place"""

    def __init__(self, prob):
        self.P = prob
        self.L = prob.L
        self.h = C.c_void_p(self.L.orc_ms_create(prob.p))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_ms_destroy(self.h)
            self.h = None

    def reset(self, dg, parameter):
        assert dg.dtype == np.int32 and dg.flags["C_CONTIGUOUS"]
        return int(self.L.orc_ms_reset(self.h, _ip(dg), int(parameter)))

    def step_stream(self, dg, parameter, us):
        assert dg.dtype == np.int32 and dg.flags["C_CONTIGUOUS"]
        us = np.ascontiguousarray(us, dtype=np.float64)
        used = C.c_int(0)
        rc = int(self.L.orc_ms_step_stream(self.h, _ip(dg), int(parameter), _dp(us), len(us), C.byref(used)))
        return rc, used.value

    def step(self, dg, parameter, seed, chain, iteration):
        assert dg.dtype == np.int32 and dg.flags["C_CONTIGUOUS"]
        return int(self.L.orc_ms_step(self.h, _ip(dg), int(parameter), C.c_uint64(seed), C.c_uint32(chain),
                                      C.c_uint64(iteration)))

    def state(self, edges=False):
        M, N = self.P.M, self.P.N
        raw = np.zeros((M, 2)); fwd = np.zeros((M, 2)); fb = np.zeros((M, 2))
        e = np.zeros((M, 2 * N), np.int32) if edges else None
        self.L.orc_ms_state(self.h, _dp(raw), _dp(fwd), _dp(fb), _ip(e) if edges else None)
        return dict(raw=raw, fwd=fwd, fb=fb, edges=e)


def elod_replicate(p1, p2, replicate, seed, chain):
    """(sampled three-locus graph int32 [3,N,2], ln-prob) of one ELOD replicate"""
    dg = np.zeros((3, p1.N, 2), np.int32)
    v = lib().orc_elod_replicate(p1.p, p2.p, C.c_long(int(replicate)), C.c_uint64(seed), C.c_uint32(chain), _ip(dg))
    return dg, float(v)


def marker_prob(isfounder, typed, genotype, xmale, mapprob):
    m = np.ascontiguousarray(mapprob, dtype=np.float64)
    out = np.zeros(4)
    lib().orc_marker_prob(int(isfounder), int(typed), int(genotype), int(xmale), _dp(m), _dp(out))
    return out


def log_sum(a, b):
    return float(lib().orc_log_sum(C.c_double(a), C.c_double(b)))


def lod_normalise(score, count, trait_prob):
    return float(lib().orc_lod_normalise(C.c_double(score), int(count), C.c_double(trait_prob)))


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*[int(x) for x in ctr])
    k = (C.c_uint32 * 2)(*[int(x) for x in key])
    o = (C.c_uint32 * 4)()
    lib().orc_philox(c, k, o)
    return [int(x) for x in o]


def uniform(seed, chain, iteration, locus, slot):
    return float(lib().orc_uniform_draw(C.c_uint64(seed), C.c_uint32(chain), C.c_uint64(iteration),
                                        C.c_uint32(locus), C.c_uint32(slot)))


def problem_from_ref(ref):
    """Builds the oracle's flat problem from a loaded reference context (oracle/refapi.Ref)
    whose peel sequence has been built."""
    pt = ref.person_table()
    mt = ref.map_table()
    return dict(N=ref.N, F=ref.F, M=ref.M, nlod=ref.nlod, sex_linked=ref.sex_linked,
                mother=pt["mother"], father=pt["father"], sex=pt["sex"], affection=pt["affection"],
                typed=pt["typed"], disease_prob=pt["disease_prob"], marker_prob=ref.marker_trait_prob(),
                genotypes=ref.genotypes(), elim=ref.elim_masks(), theta=mt["theta"], partial=mt["partial"],
                gdist=mt["gdist"], minor=mt["minor"], mapprob=mt["prob"], mapxprob=mt["xprob"],
                ops=ref.ops())
