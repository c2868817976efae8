"""The C restatement of the M-sampler (oracle/msampler_oracle.c) against golden vectors frozen
from the compiled reference (tests/golden/make_golden_ms.py) and, where oracle/_ref is present,
against the reference itself.  Bit-exact: integer edge lists, same floating-point operation order."""
import numpy as np
import pytest

from common import CASES, FORCE_X, case_files, golden, oracle_problem, problem, ref_available


def golden_ms(name):
    return golden(name + "_ms")


@pytest.mark.parametrize("name", CASES)
def test_founder_allele_graph(name):
    """FounderAlleleGraph4::reset / flip / likelihood (founder_allele_graph4.cc:34-598)"""
    fx, ms, orc = golden(name), golden_ms(name), oracle_problem(name)
    assert (orc.ms_ordering() == ms["ms_ordering"]).all()
    for gi in range(fx["dgs"].shape[0]):
        dg = fx["dgs"][gi]
        for l in range(orc.M):
            e, lik = orc.fag(dg, l)
            assert (e == ms["fag_edges_%d" % gi][l]).all()
            assert lik == ms["fag_lik_%d" % gi][l]
            for k, m in enumerate(ms["flip_meioses"]):
                _, fl = orc.fag(dg, l, (orc.F + int(m) // 2, int(m) % 2))
                assert fl == ms["fag_flip_lik_%d" % gi][l, k]


@pytest.mark.parametrize("name", CASES)
def test_descent_graph_likelihood(name):
    """DescentGraph::get_likelihood (descent_graph.cc:150-265)"""
    fx, orc = golden(name), oracle_problem(name)
    for gi in range(fx["dgs"].shape[0]):
        assert orc.dg_likelihood(fx["dgs"][gi]) == fx["dg_likelihood"][gi]


@pytest.mark.parametrize("name", CASES)
def test_msampler_trace(name):
    """MeiosisSampler::reset + step over every meiosis, fed the reference's own uniforms
    (meiosis_sampler.cc:17-203)"""
    fx, ms, orc = golden(name), golden_ms(name), oracle_problem(name)
    dg = np.ascontiguousarray(fx["dgs"][2]).copy()
    s = orc.msampler()
    assert s.reset(dg, int(ms["ms_ordering"][0])) == 0
    for k, m in enumerate(ms["trace_visit"]):
        rc, used = s.step_stream(dg, int(m), ms["trace_us"][k])
        assert rc == 0
        st = s.state()
        assert (st["raw"] == ms["trace_raw"][k]).all()
        assert (st["fb"] == ms["trace_fb"][k]).all()
        assert (dg == ms["trace_dg"][k]).all()


def test_philox_sweep_is_deterministic_and_legal():
    orc, fx = oracle_problem("east"), golden("east")
    a = np.ascontiguousarray(fx["dgs"][2]).copy()
    b = a.copy()
    assert orc.ms_sweep(a, 5, 1, 9) == 0 and orc.ms_sweep(b, 5, 1, 9) == 0
    assert (a == b).all() and (a != fx["dgs"][2]).any()
    assert orc.dg_likelihood(a) > -1e300
    c = np.ascontiguousarray(fx["dgs"][2]).copy()
    assert orc.ms_sweep(c, 5, 1, 10) == 0 and (c != a).any()
    order = orc.ms_ordering()
    sh = orc.ms_shuffle(order, 5, 1, 9)
    assert sorted(sh.tolist()) == sorted(order.tolist()) and (sh != order).any()


@pytest.mark.skipif(not ref_available(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("name", CASES)
def test_against_reference_live(name, tmp_path):
    """fresh draws, every graph: the reference and the restatement side by side"""
    from oracle import refapi as R
    R.set_threads(1)
    R.seed(4242)
    fx, orc = golden(name), oracle_problem(name)
    r = R.Ref(*case_files(name, tmp_path), sex_linked=bool(FORCE_X[name]))
    assert r.set_peel(np.array([o["peelnode"] for o in problem(name)["ops"]], np.uint32))
    order = orc.ms_ordering()
    assert (r.ms_ordering() == order).all()
    for gi in (0, 1, 3):
        dg = np.ascontiguousarray(fx["dgs"][gi]).copy()
        r.dg_set(dg)
        assert r.dg_likelihood() == orc.dg_likelihood(dg)
        s = orc.msampler()
        r.ms_reset(int(order[-1]))
        assert s.reset(dg, int(order[-1])) == 0
        for m in order[::-1]:
            us, raw, fb = r.ms_step(int(m))
            rc, _ = s.step_stream(dg, int(m), us)
            st = s.state()
            assert rc == 0 and (st["raw"] == raw).all() and (st["fb"] == fb).all()
            assert (dg == r.dg_get()).all()
    r.close()
