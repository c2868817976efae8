"""The peel kernels' per-thread code (swiftlink_b200/csrc/slk_peel.h) and the plan flattening (slk_plan.cc:
sorted-digit matrix layout, pre-decoded records, gather runs, row maps), run SEQUENTIALLY ON THE CPU by the
emulation harness tests/emu/slk_emu.cc and compared bit for bit with the C oracle and the golden vectors frozen
from the compiled reference.  Needs no GPU: it is what lets the index arithmetic be checked before it ever runs on
a device.  The emulation is test infrastructure -- the product library neither contains nor loads it."""
import ctypes as C

import numpy as np
import pytest

from common import CASES, golden, problem, oracle_problem


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


class Emu(object):
    def __init__(self, d):
        import sys, os
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
        import build as emu_build
        from swiftlink_b200 import capi
        self.L = C.CDLL(emu_build.build())
        self.L.emu_create.restype = C.c_void_p
        self.L.emu_ls_step.restype = C.c_double
        self.L.emu_lod_position.restype = C.c_double
        self.pb, self.keep = capi.make_problem(d)
        self.h = C.c_void_p(self.L.emu_create(C.byref(self.pb)))
        assert self.h, "plan rejected"
        self.N, self.M, self.nlod = int(d["N"]), int(d["M"]), int(d["nlod"])
        self.nops = len(d["ops"])
        self.cells = sum(4 ** len(o["cutset"]) for o in d["ops"])

    def close(self):
        if self.h:
            self.L.emu_destroy(self.h)
            self.h = None

    def stat(self, which):
        return int(self.L.emu_stat(self.h, which))

    def forward(self, dg, locus, ignore_left=False, ignore_right=False):
        dg = np.ascontiguousarray(dg, np.int32).copy()
        mat = np.zeros(self.cells); pre = np.zeros(4 * self.cells)
        r = self.L.emu_ls_step(self.h, _ip(dg), int(locus), int(ignore_left), int(ignore_right), 1,
                               C.c_uint64(0), C.c_uint32(0), C.c_uint64(0), _dp(mat), _dp(pre), None, None)
        return float(r), mat, pre

    def step(self, dg, locus, seed, chain, iteration, ignore_left=False, ignore_right=False):
        assert dg.dtype == np.int32 and dg.flags["C_CONTIGUOUS"]
        pmk = np.zeros(self.N, np.int32); dist = np.zeros((self.nops, 4))
        r = self.L.emu_ls_step(self.h, _ip(dg), int(locus), int(ignore_left), int(ignore_right), 0,
                               C.c_uint64(seed), C.c_uint32(chain), C.c_uint64(iteration), None, None, _ip(pmk), _dp(dist))
        return float(r), pmk, dist

    def lod_position(self, dg, interval, k, dump=False):
        dg = None if dg is None else np.ascontiguousarray(dg, np.int32)
        mat = np.zeros(self.cells) if dump else None
        prob = C.c_double(0)
        r = self.L.emu_lod_position(self.h, _ip(dg), int(interval), int(k), C.byref(prob), _dp(mat))
        return float(r), prob.value, mat


@pytest.fixture(scope="module")
def emus():
    made = {}

    def get(name):
        if name not in made:
            made[name] = Emu(problem(name))
        return made[name]
    yield get
    for e in made.values():
        e.close()


@pytest.mark.parametrize("name", CASES)
def test_forward_matrices_bit_exact(emus, name):
    """every peel matrix and presum matrix of every locus == oracle == golden vectors of the reference"""
    fx, orc, em = golden(name), oracle_problem(name), emus(name)
    for gi in range(fx["dgs"].shape[0]):
        dg = np.ascontiguousarray(fx["dgs"][gi])
        for l in range(orc.M):
            r0, m0, p0 = orc.ls_forward(dg, l)
            r1, m1, p1 = em.forward(dg, l)
            assert r0 == r1, (name, gi, l)
            assert (m0 == m1).all() and (p0 == p1).all(), (name, gi, l)
        for l in fx["sample_loci"]:
            _, m1, p1 = em.forward(dg, int(l))
            assert (m1 == fx["ls_mat_%d_%d" % (gi, l)]).all()
            assert (p1 == fx["ls_pre_%d_%d" % (gi, l)]).all()
        _, m1, _ = em.forward(dg, int(fx["sample_loci"][1]), ignore_left=True, ignore_right=False)
        assert (m1 == fx["ls_si_mat_%d" % gi]).all()


@pytest.mark.parametrize("name", CASES)
def test_steps_match_oracle(emus, name):
    """sampling 4-vectors, sampled genotypes and the written indicators, locus by locus"""
    fx, orc, em = golden(name), oracle_problem(name), emus(name)
    dg0 = np.ascontiguousarray(fx["dgs"][1]).copy()
    dg1 = dg0.copy()
    for it in range(2):
        for l in range(orc.M):
            ro, pmko, disto = orc.ls_step(dg0, l, 77, 3, it)
            rg, pmkg, distg = em.step(dg1, l, 77, 3, it)
            assert ro == rg and (pmko == pmkg).all() and (disto == distg).all(), (name, it, l)
            assert (dg0 == dg1).all()


@pytest.mark.parametrize("name", CASES)
def test_lod_positions_match_oracle(emus, name):
    fx, orc, em = golden(name), oracle_problem(name), emus(name)
    dg = np.ascontiguousarray(fx["dgs"][0])
    for interval in range(orc.M - 1):
        res, prob, mat = orc.lod_interval(dg, interval, dump_k=1 if orc.nlod > 1 else 0)
        for k in range(orc.nlod):
            r, p, m = em.lod_position(dg, interval, k, dump=(k == (1 if orc.nlod > 1 else 0)))
            assert r == res[k], (name, interval, k)
            assert abs(p - prob[k]) <= 1e-12 * abs(prob[k])
            if m is not None:
                assert (m == mat).all(), (name, interval, k)
    r, p, _ = em.lod_position(None, 0, 0)
    assert abs(p - orc.trait_prob()) <= 1e-12 * abs(p)


def test_bench_pedigree_matches_oracle():
    """the bench workload's pedigree (200 members, cutsets up to 6: matrices split between the shared-memory arena and
    the global slab, padded and plain layouts, multi-run gathers), first 40 SNPs"""
    import sys
    from common import ROOT
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench
    from oracle import orcapi
    from swiftlink_b200 import host as H
    hst = H.Host(*bench.workload_files(40, "t_emu"), lodscores=bench.N_LOD)
    assert hst.set_peel_by_names(bench.load_order()["order"])
    d = hst.problem_dict()
    orc, em = orcapi.Problem(d), Emu(d)
    assert em.stat(0) < em.stat(1), "the sampler arena is expected to spill into the global slab"
    dg0 = np.zeros((hst.M, hst.N, 2), np.int32)
    for l in range(hst.M):                                   # LocusSampler::locus_by_locus
        orc.ls_step(dg0, l, 5, 1, 0, ignore_left=True, ignore_right=True)
    dg1 = dg0.copy()
    for l in (0, 1, 17, hst.M - 1):
        r0, m0, p0 = orc.ls_forward(dg0, l)
        r1, m1, p1 = em.forward(dg0, l)
        assert r0 == r1 and (m0 == m1).all() and (p0 == p1).all(), l
    for l in range(hst.M):
        ro, pmko, disto = orc.ls_step(dg0, l, 77, 3, 9)
        rg, pmkg, distg = em.step(dg1, l, 77, 3, 9)
        assert ro == rg and (pmko == pmkg).all() and (disto == distg).all(), l
    assert (dg0 == dg1).all()
    for interval in (0, 11, hst.M - 2):
        res, prob, mat = orc.lod_interval(dg0, interval, dump_k=2)
        for k in range(orc.nlod):
            r, p, m = em.lod_position(dg0, interval, k, dump=(k == 2))
            assert r == res[k]
            if m is not None:
                assert (m == mat).all()
    em.close(); hst.close()
