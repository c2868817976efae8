"""ctypes view of include/swiftlink_b200.h.

Python here is plumbing for the tests and bench.py only: every call goes straight through the
C ABI of libswiftlink_b200.so.  There is no fallback -- if the library is missing or no sm_100
device is usable the calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SLK_LIB: an alternative build of the same library (A/B timing of kernel variants; tools/gpu/*.sh)
LIB_PATH = os.environ.get("SLK_LIB") or os.path.join(HERE, "libswiftlink_b200.so")

MAX_CUTSET, MAX_PREV, MAX_CHILDREN = 10, 8, 10

OK = 0
ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_ZERO_LIKELIHOOD, ERR_NONPOSITIVE_TRAIT, ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6

# every symbol include/swiftlink_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "slk_abi_version", "slk_last_error", "slk_device_count",
    "slk_plan_create", "slk_plan_destroy", "slk_plan_stats", "slk_plan_validate",
    "slk_chain_create", "slk_chain_destroy", "slk_chain_set_stream", "slk_chain_sync",
    "slk_dg_upload", "slk_dg_download", "slk_dg_swap",
    "slk_lsampler_window", "slk_lsampler_sweep", "slk_lsampler_locus_by_locus", "slk_sequential_imputation", "slk_sequential_imputation_batch",
    "slk_lodscore_init", "slk_lodscore_accumulate", "slk_lodscore_read", "slk_lodscore_normalise",
    "slk_trait_likelihood", "slk_elod_run", "slk_debug_elod_graphs",
    "slk_msampler_ordering", "slk_msampler_reset", "slk_msampler_step", "slk_msampler_sweep", "slk_dg_likelihood",
    "slk_sweep_is_lsampler", "slk_debug_fag", "slk_debug_msampler_state", "slk_debug_msampler_trace", "slk_debug_msampler_launch", "slk_debug_msampler_timeline",
    "slk_debug_lsampler_forward", "slk_debug_lsampler_step", "slk_debug_lod_interval", "slk_debug_lsampler_trace",
    "slk_debug_philox", "slk_debug_uniform", "slk_measure_fp64_peak",
]


class PeelOp(C.Structure):
    _fields_ = [("type", C.c_int32), ("peelnode", C.c_int32),
                ("ncut", C.c_int32), ("cutset", C.c_int32 * MAX_CUTSET),
                ("nprev", C.c_int32), ("prev", C.c_int32 * MAX_PREV),
                ("nchild", C.c_int32), ("children", C.c_int32 * MAX_CHILDREN)]


class Problem(C.Structure):
    _fields_ = [("n_members", C.c_int32), ("n_founders", C.c_int32), ("n_markers", C.c_int32),
                ("n_lod", C.c_int32), ("sex_linked", C.c_int32),
                ("mother", C.POINTER(C.c_int32)), ("father", C.POINTER(C.c_int32)),
                ("sex", C.POINTER(C.c_int32)), ("typed", C.POINTER(C.c_int32)),
                ("prior_as_founder", C.POINTER(C.c_int32)),
                ("genotypes", C.POINTER(C.c_uint8)), ("disease_prob", C.POINTER(C.c_double)),
                ("marker_prob", C.POINTER(C.c_double)), ("marker_xprob", C.POINTER(C.c_double)),
                ("theta", C.POINTER(C.c_double)), ("partial_theta", C.POINTER(C.c_double)),
                ("elimination", C.POINTER(C.c_uint8)),
                ("n_ops", C.c_int32), ("ops", C.POINTER(PeelOp)),
                ("minor_freq", C.POINTER(C.c_double)), ("disease_prior_locus_plus1", C.c_int32),
                ("person_prior", C.POINTER(C.c_double))]


class SlkError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "swiftlink_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run `python -m swiftlink_b200.build` (there is no CPU fallback)"
                               % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.slk_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _check(rc):
    if rc != OK:
        raise SlkError(rc, lib().slk_last_error().decode())


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def device_count():
    return int(lib().slk_device_count())


def make_problem(d):
    """dict of flat arrays (the layout of oracle.orcapi.problem_from_ref / swiftlink_b200.host)
    -> (Problem struct, keep-alive list)."""
    keep = {}
    N, M = int(d["N"]), int(d["M"])
    keep["mother"] = np.ascontiguousarray(d["mother"], np.int32)
    keep["father"] = np.ascontiguousarray(d["father"], np.int32)
    keep["sex"] = np.ascontiguousarray(d["sex"], np.int32)
    keep["typed"] = np.ascontiguousarray(d["typed"], np.int32)
    paf = d.get("prior_as_founder")
    keep["paf"] = None if paf is None else np.ascontiguousarray(paf, np.int32)
    keep["genotypes"] = np.ascontiguousarray(d["genotypes"], np.uint8).reshape(N, M)
    keep["disease_prob"] = np.ascontiguousarray(d["disease_prob"], np.float64)
    keep["mapprob"] = np.ascontiguousarray(d["mapprob"], np.float64)
    keep["mapxprob"] = np.ascontiguousarray(d["mapxprob"], np.float64)
    keep["theta"] = np.ascontiguousarray(d["theta"], np.float64)
    keep["partial"] = np.ascontiguousarray(d["partial"], np.float64)
    keep["elim"] = np.ascontiguousarray(d["elim"], np.uint8).reshape(M, N)
    ops = d["ops"]
    arr = (PeelOp * len(ops))()
    for i, o in enumerate(ops):
        arr[i].type = int(o["type"])
        arr[i].peelnode = int(o["peelnode"])
        for name, cnt, key, cap in (("cutset", "ncut", "cutset", MAX_CUTSET), ("prev", "nprev", "previous", MAX_PREV),
                                    ("children", "nchild", "children", MAX_CHILDREN)):
            vals = [int(x) for x in o[key]]
            if len(vals) > cap:
                raise ValueError("op %d: %s has %d entries (max %d)" % (i, key, len(vals), cap))
            setattr(arr[i], cnt, len(vals))
            for j, v in enumerate(vals):
                getattr(arr[i], name)[j] = v
    keep["ops"] = arr
    keep["minor"] = None if d.get("minor") is None else np.ascontiguousarray(d["minor"], np.float64)
    keep["person_prior"] = None if d.get("person_prior") is None else np.ascontiguousarray(d["person_prior"], np.float64)
    p = Problem(N, int(d["F"]), M, int(d["nlod"]), int(d["sex_linked"]),
                _ptr(keep["mother"], C.c_int32), _ptr(keep["father"], C.c_int32),
                _ptr(keep["sex"], C.c_int32), _ptr(keep["typed"], C.c_int32),
                _ptr(keep["paf"], C.c_int32) if keep["paf"] is not None else None,
                _ptr(keep["genotypes"], C.c_uint8), _ptr(keep["disease_prob"], C.c_double),
                _ptr(keep["mapprob"], C.c_double), _ptr(keep["mapxprob"], C.c_double),
                _ptr(keep["theta"], C.c_double), _ptr(keep["partial"], C.c_double),
                _ptr(keep["elim"], C.c_uint8), len(ops), arr,
                _ptr(keep["minor"], C.c_double) if keep["minor"] is not None else None,
                int(d.get("disease_prior_locus", -1)) + 1,
                _ptr(keep["person_prior"], C.c_double) if keep["person_prior"] is not None else None)
    return p, keep


STAT_NAMES = ["n_ops", "sum_cells", "sum_presum", "flops_ls", "flops_lod", "ls_flevels", "ls_blevels",
              "lod_flevels", "ls_arena_doubles", "lod_arena_doubles", "lod_valid_cells", "max_cutset",
              "ls_team_threads", "lod_team_threads", "ls_smem_doubles", "lod_smem_doubles",
              "ls_blocks_per_sm", "lod_blocks_per_sm", "ls_cta_smem", "lod_cta_smem"]


def plan_validate(d):
    """host-only flattening of a problem; returns the plan statistics (no device needed)"""
    prob, keep = make_problem(d)
    out = np.zeros(len(STAT_NAMES))
    _check(lib().slk_plan_validate(C.byref(prob), _ptr(out, C.c_double), len(STAT_NAMES)))
    return dict(zip(STAT_NAMES, out.tolist()))


class Plan(object):
    def __init__(self, d, device=0):
        self.L = lib()
        self.N, self.F, self.M, self.nlod = int(d["N"]), int(d["F"]), int(d["M"]), int(d["nlod"])
        self.nops = len(d["ops"])
        self.sum_cells = sum(4 ** len(o["cutset"]) for o in d["ops"])
        self.sum_presum = 4 * self.sum_cells
        prob, keep = make_problem(d)
        self.h = C.c_void_p()
        _check(self.L.slk_plan_create(C.byref(prob), int(device), C.byref(self.h)))
        self.device = device

    def stats(self):
        out = np.zeros(len(STAT_NAMES))
        n = self.L.slk_plan_stats(self.h, _ptr(out, C.c_double), len(STAT_NAMES))
        return dict(zip(STAT_NAMES[:n], out[:n].tolist()))

    def trait_likelihood(self):
        v = C.c_double(0)
        _check(self.L.slk_trait_likelihood(self.h, C.byref(v)))
        return v.value

    def msampler_ordering(self):
        out = np.zeros(2 * self.N + 1, np.int32)
        n = int(self.L.slk_msampler_ordering(self.h, _ptr(out, C.c_int32), len(out)))
        return out[:n].copy()

    def close(self):
        if self.h:
            self.L.slk_plan_destroy(self.h)
            self.h = None


class Chain(object):
    def __init__(self, plan, seed=1, chain_id=0):
        self.L = lib()
        self.plan = plan
        self.h = C.c_void_p()
        _check(self.L.slk_chain_create(plan.h, C.c_uint64(seed), C.c_uint32(chain_id), C.byref(self.h)))

    def close(self):
        if self.h:
            self.L.slk_chain_destroy(self.h)
            self.h = None

    def set_stream(self, cuda_stream_ptr):
        _check(self.L.slk_chain_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def sync(self):
        _check(self.L.slk_chain_sync(self.h))

    def dg_upload(self, dg):
        dg = np.ascontiguousarray(dg, np.int32)
        assert dg.shape == (self.plan.M, self.plan.N, 2)
        _check(self.L.slk_dg_upload(self.h, _ptr(dg, C.c_int32)))

    def dg_upload_ptr(self, ptr):
        _check(self.L.slk_dg_upload(self.h, C.c_void_p(ptr)))

    def dg_download(self, out=None):
        if out is None:
            out = np.zeros((self.plan.M, self.plan.N, 2), np.int32)
        _check(self.L.slk_dg_download(self.h, _ptr(out, C.c_int32)))
        return out

    def dg_download_ptr(self, ptr):
        _check(self.L.slk_dg_download(self.h, C.c_void_p(ptr)))

    def dg_swap(self, other):
        _check(self.L.slk_dg_swap(self.h, other.h))

    def lsampler_window(self, iteration, window, offset):
        _check(self.L.slk_lsampler_window(self.h, C.c_uint64(iteration), int(window), int(offset)))

    def lsampler_sweep(self, iteration):
        _check(self.L.slk_lsampler_sweep(self.h, C.c_uint64(iteration)))

    def lsampler_locus_by_locus(self, iteration=0):
        _check(self.L.slk_lsampler_locus_by_locus(self.h, C.c_uint64(iteration)))

    def sequential_imputation(self, run=0, start_locus=0, want_weight=True):
        w = C.c_double(0)
        _check(self.L.slk_sequential_imputation(self.h, C.c_uint64(run), int(start_locus),
                                                C.byref(w) if want_weight else None))
        return w.value if want_weight else None

    def sequential_imputation_batch(self, start_loci, first_run=0):
        """n independent start_from walks in one launch per wave of teams; returns (log weights, index of the best),
        the best walk's graph becomes the chain's graph"""
        st = np.ascontiguousarray(start_loci, np.int32)
        w = np.zeros(len(st))
        best = C.c_int32(-1)
        _check(self.L.slk_sequential_imputation_batch(self.h, C.c_uint64(first_run), int(len(st)),
                                                      st.ctypes.data_as(C.POINTER(C.c_int32)),
                                                      w.ctypes.data_as(C.POINTER(C.c_double)), C.byref(best)))
        return w, int(best.value)

    def lodscore_init(self):
        _check(self.L.slk_lodscore_init(self.h))

    def lodscore_accumulate(self):
        _check(self.L.slk_lodscore_accumulate(self.h))

    def lodscore_read(self, out=None):
        n = (self.plan.M - 1) * self.plan.nlod
        if out is None:
            out = np.zeros(n)
        cnt = C.c_int32(0)
        _check(self.L.slk_lodscore_read(self.h, _ptr(out, C.c_double), C.byref(cnt)))
        return out.reshape(self.plan.M - 1, self.plan.nlod), cnt.value

    def lodscore_read_ptr(self, ptr):
        cnt = C.c_int32(0)
        _check(self.L.slk_lodscore_read(self.h, C.c_void_p(ptr), C.byref(cnt)))
        return cnt.value

    def lodscore_normalise(self, trait_prob):
        out = np.zeros((self.plan.M - 1) * self.plan.nlod)
        _check(self.L.slk_lodscore_normalise(self.h, C.c_double(trait_prob), _ptr(out, C.c_double)))
        return out.reshape(self.plan.M - 1, self.plan.nlod)

    # ---- M-sampler ------------------------------------------------------------------------
    def msampler_reset(self):
        _check(self.L.slk_msampler_reset(self.h))

    def msampler_step(self, iteration, meiosis):
        _check(self.L.slk_msampler_step(self.h, C.c_uint64(iteration), int(meiosis)))

    def msampler_sweep(self, iteration):
        _check(self.L.slk_msampler_sweep(self.h, C.c_uint64(iteration)))

    def dg_likelihood(self):
        v = C.c_double(0)
        _check(self.L.slk_dg_likelihood(self.h, C.byref(v)))
        return v.value

    def sweep_is_lsampler(self, iteration, lsampler_prob):
        return bool(self.L.slk_sweep_is_lsampler(self.h, C.c_uint64(iteration), C.c_double(lsampler_prob)))

    def debug_fag(self, meiosis=-1, edges=False):
        lik = np.zeros(self.plan.M)
        e = np.zeros((self.plan.M, 2 * self.plan.N), np.int32) if edges else None
        _check(self.L.slk_debug_fag(self.h, int(meiosis), _ptr(lik, C.c_double),
                                    _ptr(e, C.c_int32) if edges else None))
        return lik, e

    def debug_msampler_trace(self, m0, m1):
        buf = np.zeros((20, 8), np.int64)
        _check(self.L.slk_debug_msampler_trace(self.h, int(m0), int(m1), _ptr(buf, C.c_longlong)))
        return buf

    def debug_msampler_timeline(self, iteration, cta_pair=100):
        """%globaltimer stamps of one M-sweep: (launch stamps [n + 2][8], CTA stamps [3 * grid][2]) -- see the header"""
        n = len(self.plan.msampler_ordering())
        grid = (self.plan.M + 31) // 32
        buf = np.zeros(8 * (n + 2) + 8 * grid, np.uint64)
        _check(self.L.slk_debug_msampler_timeline(self.h, C.c_uint64(iteration), int(cta_pair), _ptr(buf, C.c_ulonglong), int(buf.size)))
        return buf[:8 * (n + 2)].reshape(n + 2, 8), buf[8 * (n + 2):].reshape(4 * grid, 2)

    def debug_msampler_launch(self, m0, m1, which, reps):
        _check(self.L.slk_debug_msampler_launch(self.h, int(m0), int(m1), int(which), int(reps)))

    def debug_msampler_state(self):
        fb = np.zeros((self.plan.M, 2)); cur = np.zeros(self.plan.M)
        _check(self.L.slk_debug_msampler_state(self.h, _ptr(fb, C.c_double), _ptr(cur, C.c_double)))
        return fb, cur

    # ---- parity hooks ---------------------------------------------------------------------
    def debug_forward(self, locus, ignore_left=False, ignore_right=False):
        mat = np.zeros(self.plan.sum_cells); pre = np.zeros(self.plan.sum_presum); res = C.c_double(0)
        _check(self.L.slk_debug_lsampler_forward(self.h, int(locus), int(ignore_left), int(ignore_right),
                                                 _ptr(mat, C.c_double), _ptr(pre, C.c_double), C.byref(res)))
        return res.value, mat, pre

    def debug_step(self, iteration, locus, ignore_left=False, ignore_right=False):
        pmk = np.zeros(self.plan.N, np.int32); dist = np.zeros((self.plan.nops, 4)); res = C.c_double(0)
        _check(self.L.slk_debug_lsampler_step(self.h, C.c_uint64(iteration), int(locus), int(ignore_left),
                                              int(ignore_right), _ptr(pmk, C.c_int32), _ptr(dist, C.c_double),
                                              C.byref(res)))
        return res.value, pmk, dist

    def debug_trace(self, iteration=0, offset=0):
        buf = (C.c_longlong * 512)()
        n = self.L.slk_debug_lsampler_trace(self.h, C.c_uint64(iteration), int(offset), buf, 512)
        if n < 0:
            _check(n)
        return np.array(buf[:n], dtype=np.int64)

    def debug_lod_interval(self, interval, dump_k=-1):
        res = np.zeros(self.plan.nlod); prob = np.zeros(self.plan.nlod)
        mat = np.zeros(self.plan.sum_cells) if dump_k >= 0 else None
        _check(self.L.slk_debug_lod_interval(self.h, int(interval), _ptr(res, C.c_double), _ptr(prob, C.c_double),
                                             int(dump_k), _ptr(mat, C.c_double) if mat is not None else None))
        return res, prob, mat


def elod_run(sampler_plan, trait_plan, replicates, seed=1, chain_id=0, want_probs=False):
    """Elod::run's replicate loop on the device: returns (log-sum accumulator, count, per-replicate ln-probs or None)"""
    ls = C.c_double(0); cnt = C.c_int64(0)
    probs = np.zeros(int(replicates)) if want_probs else None
    _check(lib().slk_elod_run(sampler_plan.h, trait_plan.h, C.c_uint64(seed), C.c_uint32(chain_id), C.c_int64(int(replicates)),
                              C.byref(ls), C.byref(cnt), _ptr(probs, C.c_double) if want_probs else None))
    return ls.value, cnt.value, probs


def debug_elod_graphs(sampler_plan, first, n, seed=1, chain_id=0):
    dg = np.zeros((int(n), 3, sampler_plan.N, 2), np.int32)
    _check(lib().slk_debug_elod_graphs(sampler_plan.h, C.c_uint64(seed), C.c_uint32(chain_id), C.c_int64(int(first)), int(n),
                                       _ptr(dg, C.c_int32)))
    return dg


def debug_philox(ctr, key, device=0):
    c = (C.c_uint32 * 4)(*[int(x) for x in ctr])
    k = (C.c_uint32 * 2)(*[int(x) for x in key])
    o = (C.c_uint32 * 4)()
    _check(lib().slk_debug_philox(int(device), c, k, o))
    return [int(x) for x in o]


def debug_uniform(seed, chain, iteration, locus, slot, device=0):
    v = C.c_double(0)
    _check(lib().slk_debug_uniform(int(device), C.c_uint64(seed), C.c_uint32(chain), C.c_uint64(iteration),
                                   C.c_uint32(locus), C.c_uint32(slot), C.byref(v)))
    return v.value


def measure_fp64_peak(device=0):
    v = C.c_double(0)
    _check(lib().slk_measure_fp64_peak(int(device), C.byref(v)))
    return v.value
