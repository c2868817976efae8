"""The C oracle (oracle/peel_oracle.c) against the golden vectors frozen from the compiled
reference (tests/golden/make_golden.py).  Bit-exact: same operation order, no FMA."""
import numpy as np
import pytest

from common import CASES, golden, problem, oracle_problem, unpack_ops
from oracle import orcapi


@pytest.mark.parametrize("name", CASES)
def test_index_tables(name):
    fx, orc = golden(name), oracle_problem(name)
    for i in range(orc.nops):
        assert (orc.op_indices(i, 0) == fx["lod_indices_%d" % i]).all()
        for l in fx["sample_loci"]:
            assert (orc.op_indices(i, 1, int(l)) == fx["matrix_indices_%d_%d" % (i, l)]).all()
            assert (orc.op_indices(i, 2, int(l)) == fx["presum_indices_%d_%d" % (i, l)]).all()


@pytest.mark.parametrize("name", CASES)
def test_sampler_forward_matrices(name):
    fx, orc = golden(name), oracle_problem(name)
    for gi in range(fx["dgs"].shape[0]):
        dg = fx["dgs"][gi]
        for l in fx["sample_loci"]:
            res, mat, pre = orc.ls_forward(dg, int(l))
            assert res == float(fx["ls_result_%d_%d" % (gi, l)])
            assert (mat == fx["ls_mat_%d_%d" % (gi, l)]).all()
            assert (pre == fx["ls_pre_%d_%d" % (gi, l)]).all()
        l = int(fx["sample_loci"][1])
        res, mat, pre = orc.ls_forward(dg, l, ignore_left=True, ignore_right=False)
        assert (mat == fx["ls_si_mat_%d" % gi]).all()


@pytest.mark.parametrize("name", CASES)
def test_lod_scoring(name):
    fx, orc = golden(name), oracle_problem(name)
    assert orc.trait_prob() == float(fx["trait_prob"])
    assert orc.marker_transmission() == float(fx["marker_transmission"])
    for gi in range(fx["dgs"].shape[0]):
        dg = fx["dgs"][gi]
        for itv in fx["sample_intervals"]:
            res, prob, mat = orc.lod_interval(dg, int(itv), 2)
            assert (res == fx["lod_result_%d_%d" % (gi, itv)]).all()
            assert (prob == fx["lod_prob_%d_%d" % (gi, itv)]).all()
            assert (mat == fx["lod_mat_%d_%d" % (gi, itv)]).all()
        rec = np.array([orc.recombination_prob(dg, l) for l in range(orc.M - 1)])
        assert (rec == fx["recomb_%d" % gi]).all()
    # one full pass (peeler.cc:79-103 over every interval) on the last graph
    gi = fx["dgs"].shape[0] - 1
    sc = np.zeros((orc.M - 1) * orc.nlod)
    orc.lod_pass(fx["dgs"][gi], sc, True)
    assert (sc.reshape(orc.M - 1, orc.nlod) == fx["lod_pass_%d" % gi]).all()


@pytest.mark.parametrize("name", CASES)
def test_marker_prior(name):
    """Person::populate_trait_prob_cache restated from genotype codes and map probabilities.

    The reference fills this cache while it parses the ped file (pedigree_parser.cc:153), before
    parent ids are resolved, so Person::isfounder() is true for EVERY person at that point and
    untyped non-founders get the population prior as well.  The `inbred` fixture (allele
    frequencies != 0.5, untyped non-founders) is the case that tells the two readings apart."""
    fx = golden(name)
    N, M, F, X = int(fx["N"]), int(fx["M"]), int(fx["F"]), int(fx["sex_linked"])
    for i in range(N):
        xmale = bool(X and fx["sex"][i] == 1)
        for l in range(M):
            mp = fx["mapxprob"][l] if xmale else fx["mapprob"][l]
            got = orcapi.marker_prob(True, fx["typed"][i], fx["genotypes"][i, l], xmale, mp)
            assert (got == fx["marker_prob"][i, l]).all(), (i, l)


def test_log_sum_and_normalise():
    a, b = -31.2, -33.9
    assert orcapi.log_sum(a, b) == np.log(np.exp(b - a) + 1) + a
    big = -np.finfo(np.float64).max
    assert orcapi.log_sum(big, b) == b and orcapi.log_sum(a, big) == a
    assert orcapi.lod_normalise(-20.0, 10, -21.0) == (-20.0 - np.log(10.0) - (-21.0)) / np.log(10.0)


@pytest.mark.parametrize("name", CASES)
def test_step_is_deterministic_and_legal(name):
    """Philox-keyed step: same key -> same graph; sampled genotypes respect the elimination masks."""
    fx, orc = golden(name), oracle_problem(name)
    dg1 = np.ascontiguousarray(fx["dgs"][1]).copy()
    dg2 = dg1.copy()
    legal_bit = {0: 8, 1: 1, 2: 2, 3: 4}      # trait code -> elimination bit (genotype.cc:91-102)
    for it in range(2):
        for l in range(orc.M):
            r1, pmk1, d1 = orc.ls_step(dg1, l, 5, 1, it)
            r2, pmk2, d2 = orc.ls_step(dg2, l, 5, 1, it)
            assert r1 == r2 and (pmk1 == pmk2).all() and (d1 == d2).all()
            for i in range(orc.N):
                assert fx["elim"][l, i] & legal_bit[int(pmk1[i])]
    assert (dg1 == dg2).all()
    assert set(np.unique(dg1)) <= {0, 1}
    assert (dg1[:, :orc.F, :] == 0).all()
