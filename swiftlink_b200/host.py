"""ctypes view of include/swiftlink_b200_host.h: the C++ host side (parsers, tables, genotype
elimination, peel-sequence generator).  Plumbing for tests and bench.py; the logic is C++."""
import ctypes as C

import numpy as np

from . import capi

SYMBOLS = [
    "slk_host_open", "slk_host_close", "slk_host_dims", "slk_host_person_table", "slk_host_person_name",
    "slk_host_marker_name", "slk_host_genotypes", "slk_host_marker_trait_prob", "slk_host_map_table",
    "slk_host_disease_model", "slk_host_elim_masks", "slk_host_build_peel", "slk_host_set_peel",
    "slk_host_num_ops", "slk_host_peel_cost", "slk_host_op_info", "slk_host_random_descentgraph",
    "slk_host_problem", "slk_host_write_results", "slk_host_run_chain", "slk_host_run_replicates", "slk_host_run_mc3", "slk_host_mc3_temperature", "slk_host_elod",
    "slk_host_job_create", "slk_host_job_advance", "slk_host_job_results", "slk_host_job_destroy",
]


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Host(object):
    """One pedigree loaded from LINKAGE ped/map/dat files by the C++ host code."""

    def __init__(self, ped, mapf, dat, sex_linked=False, lodscores=5):
        self.L = capi.lib()
        self.L.slk_host_open.restype = C.c_void_p
        self.L.slk_host_problem.restype = C.POINTER(capi.Problem)
        self.L.slk_host_peel_cost.restype = C.c_uint32
        h = self.L.slk_host_open(ped.encode(), mapf.encode(), dat.encode(), int(sex_linked), int(lodscores))
        if not h:
            raise RuntimeError("could not load %s / %s / %s" % (ped, mapf, dat))
        self.h = C.c_void_p(h)
        d = np.zeros(5, np.int32)
        self.L.slk_host_dims(self.h, _ip(d))
        self.N, self.F, self.M, self.nlod, self.sex_linked = [int(x) for x in d]

    def close(self):
        if self.h:
            self.L.slk_host_close(self.h)
            self.h = None

    def person_table(self):
        N = self.N
        mother = np.zeros(N, np.int32); father = np.zeros(N, np.int32); sex = np.zeros(N, np.int32)
        aff = np.zeros(N, np.int32); typed = np.zeros(N, np.int32); dprob = np.zeros((N, 4))
        self.L.slk_host_person_table(self.h, _ip(mother), _ip(father), _ip(sex), _ip(aff), _ip(typed), _dp(dprob))
        return dict(mother=mother, father=father, sex=sex, affection=aff, typed=typed, disease_prob=dprob)

    def _names(self, fn, n):
        out = []
        buf = C.create_string_buffer(256)
        for i in range(n):
            fn(self.h, i, buf, 256)
            out.append(buf.value.decode())
        return out

    def person_names(self):
        return self._names(self.L.slk_host_person_name, self.N)

    def marker_names(self):
        return self._names(self.L.slk_host_marker_name, self.M)

    def genotypes(self):
        g = np.zeros((self.N, self.M), np.int32)
        self.L.slk_host_genotypes(self.h, _ip(g))
        return g

    def marker_trait_prob(self):
        t = np.zeros((self.N, self.M, 4))
        self.L.slk_host_marker_trait_prob(self.h, _dp(t))
        return t

    def map_table(self):
        M = self.M
        gdist = np.zeros(M); minor = np.zeros(M); prob = np.zeros((M, 4)); xprob = np.zeros((M, 4))
        theta = np.zeros(M - 1); partial = np.zeros(M - 1)
        self.L.slk_host_map_table(self.h, _dp(gdist), _dp(minor), _dp(prob), _dp(xprob), _dp(theta), _dp(partial))
        return dict(gdist=gdist, minor=minor, prob=prob, xprob=xprob, theta=theta, partial=partial)

    def disease_model(self):
        d = np.zeros(4)
        self.L.slk_host_disease_model(self.h, _dp(d))
        return dict(freq=float(d[0]), penetrance=d[1:].copy())

    def elim_masks(self):
        m = np.zeros((self.M, self.N), np.int32)
        self.L.slk_host_elim_masks(self.h, _ip(m))
        return m

    def build_peel(self, iterations=1000000, seed=20261017):
        self.L.slk_host_build_peel(self.h, int(iterations), C.c_uint64(seed))

    def set_peel(self, seq):
        s = np.ascontiguousarray(seq, np.uint32)
        assert s.shape == (self.N,)
        return bool(self.L.slk_host_set_peel(self.h, s.ctypes.data_as(C.POINTER(C.c_uint32))))

    def set_peel_by_names(self, names):
        """elimination order given as person ids of the ped file (robust to member re-ordering)"""
        index = dict((n, i) for i, n in enumerate(self.person_names()))
        return self.set_peel([index[str(n)] for n in names])

    def peel_cost(self):
        return int(self.L.slk_host_peel_cost(self.h))

    def ops(self):
        out = []
        N = self.N
        for i in range(int(self.L.slk_host_num_ops(self.h))):
            info = np.zeros(5, np.int32); cut = np.zeros(N, np.int32); prev = np.zeros(N, np.int32)
            kids = np.zeros(N, np.int32)
            self.L.slk_host_op_info(self.h, i, _ip(info), _ip(cut), _ip(prev), _ip(kids))
            out.append(dict(type=int(info[0]), peelnode=int(info[1]), cutset=cut[:info[2]].tolist(),
                            previous=prev[:info[3]].tolist(), children=kids[:info[4]].tolist()))
        return out

    def random_descentgraph(self, seed=1):
        dg = np.zeros((self.M, self.N, 2), np.int32)
        if not self.L.slk_host_random_descentgraph(self.h, C.c_uint64(seed), _ip(dg)):
            raise RuntimeError("genotype elimination failed: inconsistent genotypes")
        return dg

    def problem_ptr(self):
        p = self.L.slk_host_problem(self.h)
        if not p:
            raise RuntimeError("peel sequence not built")
        return p

    def problem_dict(self):
        """the same layout as oracle.orcapi.problem_from_ref, produced by the host code"""
        pt = self.person_table(); mt = self.map_table()
        return dict(N=self.N, F=self.F, M=self.M, nlod=self.nlod, sex_linked=self.sex_linked,
                    mother=pt["mother"], father=pt["father"], sex=pt["sex"], affection=pt["affection"],
                    typed=pt["typed"], disease_prob=pt["disease_prob"], marker_prob=self.marker_trait_prob(),
                    genotypes=self.genotypes(), elim=self.elim_masks(), theta=mt["theta"], partial=mt["partial"],
                    gdist=mt["gdist"], minor=mt["minor"], mapprob=mt["prob"], mapxprob=mt["xprob"], ops=self.ops())

    def write_results(self, filename, lod):
        lod = np.ascontiguousarray(lod, np.float64)
        return bool(self.L.slk_host_write_results(self.h, filename.encode(), _dp(lod)))

    def run_chain(self, dg, burnin, iterations, scoring_period=10, seed=1, chain_id=0, device=0, lsampler_prob=1.0):
        dg = np.ascontiguousarray(dg, np.int32).copy()
        lod = np.zeros((self.M - 1) * self.nlod); tp = C.c_double(0)
        rc = self.L.slk_host_run_chain(self.h, int(device), C.c_uint64(seed), C.c_uint32(chain_id), int(burnin),
                                       int(iterations), int(scoring_period), C.c_double(lsampler_prob), _ip(dg), _dp(lod),
                                       C.byref(tp))
        if rc != 0:
            raise capi.SlkError(rc, self.L.slk_last_error().decode())
        return dict(lod=lod.reshape(self.M - 1, self.nlod), dg=dg, trait_prob=tp.value)


    def run_replicates(self, runs, in_flight, burnin, iterations, scoring_period=10, seed=1, device=0, lsampler_prob=0.5,
                       si_iterations=2):
        """LinkageProgram::run_pedigree's -R loop with up to `in_flight` replicate chains resident on the device at once;
        returns the merged, normalised LOD table"""
        lod = np.zeros((self.M - 1) * self.nlod)
        rc = self.L.slk_host_run_replicates(self.h, int(device), C.c_uint64(seed), int(runs), int(in_flight), int(burnin),
                                            int(iterations), int(scoring_period), C.c_double(lsampler_prob),
                                            int(si_iterations), _dp(lod))
        if rc != 0:
            raise capi.SlkError(rc, self.L.slk_last_error().decode())
        return lod.reshape(self.M - 1, self.nlod)

    def run_mc3(self, n_chains, burnin, iterations, exchange_period=10, temperatures=None, scoring_period=10,
                seed=1, chain_id=0, device=0, lsampler_prob=0.5, si_iterations=10):
        """Mc3::run: a ladder of heated chains on one device; returns the cold chain's LOD table and the swap counts"""
        lod = np.zeros((self.M - 1) * self.nlod)
        ok = np.zeros(n_chains, np.int32); bad = np.zeros(n_chains, np.int32)
        t = None if temperatures is None else np.ascontiguousarray(temperatures, np.float64)
        rc = self.L.slk_host_run_mc3(self.h, int(device), C.c_uint64(seed), C.c_uint32(chain_id), int(n_chains),
                                     int(exchange_period), _dp(t) if t is not None else None, int(burnin), int(iterations),
                                     int(scoring_period), C.c_double(lsampler_prob), int(si_iterations), _dp(lod),
                                     _ip(ok), _ip(bad))
        if rc != 0:
            raise capi.SlkError(rc, self.L.slk_last_error().decode())
        return dict(lod=lod.reshape(self.M - 1, self.nlod), swap_success=ok[:n_chains - 1], swap_failure=bad[:n_chains - 1])


class Job(object):
    """One device's share of a `-R` job (swiftlink::ReplicateJob, csrc/host/job.cc): the given replicates -- plain chains or
    MC3 ladders of `mc3_chains` chains -- resident on `device` at once.  advance(n) runs the next n iterations of every
    replicate; results() returns the RAW merged log-sum accumulators, the count of scoring passes, ln P(T) and the summed
    swap counters, ready for the cross-rank merge of swiftlink_b200.dist."""

    def __init__(self, host, replicate_ids, burnin, iterations, scoring_period=10, seed=1, device=0, lsampler_prob=0.5,
                 si_iterations=10, mc3_chains=1, exchange_period=10, temperatures=None):
        self.host, self.L = host, host.L
        self.L.slk_host_job_create.restype = C.c_void_p
        ids = np.ascontiguousarray(replicate_ids, np.int32)
        t = None if temperatures is None else np.ascontiguousarray(temperatures, np.float64)
        self.mc3_chains = max(int(mc3_chains), 1)
        h = self.L.slk_host_job_create(host.h, int(device), C.c_uint64(seed), _ip(ids), int(len(ids)), int(mc3_chains),
                                       int(exchange_period), _dp(t) if t is not None else None, int(burnin), int(iterations),
                                       int(scoring_period), C.c_double(lsampler_prob), int(si_iterations))
        if not h:
            raise RuntimeError("slk_host_job_create failed")
        self.h = C.c_void_p(h)

    def advance(self, n):
        return int(self.L.slk_host_job_advance(self.h, int(n)))

    def results(self):
        n = (self.host.M - 1) * self.host.nlod
        raw = np.zeros(n); cnt = C.c_int32(0); tp = C.c_double(0)
        ok = np.zeros(self.mc3_chains, np.int32); bad = np.zeros(self.mc3_chains, np.int32)
        rc = self.L.slk_host_job_results(self.h, _dp(raw), C.byref(cnt), C.byref(tp), _ip(ok), _ip(bad), int(self.mc3_chains))
        if rc != 0:
            raise capi.SlkError(rc, self.L.slk_last_error().decode())
        return dict(raw=raw, count=int(cnt.value), trait_prob=float(tp.value), swap_success=ok, swap_failure=bad)

    def close(self):
        if self.h:
            self.L.slk_host_job_destroy(self.h)
            self.h = None


def elod(pedfile, frequency=1e-4, penetrance=(0.0, 0.0, 1.0), separation=0.05, replicates=1000000, sex_linked=False,
         affected_only=False, peel_iterations=100000, seed=20261017, device=0):
    """Elod(pedfile, options).run() on the device: (total ELOD, per-pedigree values)"""
    L = capi.lib()
    L.slk_host_elod.restype = C.c_double
    pen = np.ascontiguousarray(penetrance, np.float64)
    per = np.zeros(64)
    total = L.slk_host_elod(pedfile.encode(), C.c_double(frequency), _dp(pen), C.c_double(separation), int(replicates),
                            int(sex_linked), int(affected_only), int(peel_iterations), C.c_uint64(seed), int(device),
                            _dp(per), 64)
    return float(total), per


def mc3_temperature(i, n_chains, temperatures=None):
    L = capi.lib()
    L.slk_host_mc3_temperature.restype = C.c_double
    t = None if temperatures is None else np.ascontiguousarray(temperatures, np.float64)
    return float(L.slk_host_mc3_temperature(int(i), int(n_chains), _dp(t) if t is not None else None))


class PlanFromHost(capi.Plan):
    """device plan built straight from the host's flattened slk_problem (no Python copies)"""

    def __init__(self, host, device=0):
        self.L = capi.lib()
        self.N, self.F, self.M, self.nlod = host.N, host.F, host.M, host.nlod
        ops = host.ops()
        self.nops = len(ops)
        self.sum_cells = sum(4 ** len(o["cutset"]) for o in ops)
        self.sum_presum = 4 * self.sum_cells
        self.h = C.c_void_p()
        capi._check(self.L.slk_plan_create(host.problem_ptr(), int(device), C.byref(self.h)))
        self.device = device
