"""ELOD on the device (slk_elod_run, swiftlink::Elod) against the C oracle replicate by replicate and against the
reference's own Elod::run (golden estimates, tests/golden/make_golden_elod.py) within Monte Carlo error."""
import os

import numpy as np
import pytest

from common import GOLDEN, elod_problems, ref_available

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["loop", "east", "inbred"])
def test_replicates_match_oracle(name):
    """every replicate's sampled three-locus graph bit-exact, its ln-probability within 1e-12, whatever the chunking"""
    from oracle import orcapi
    from swiftlink_b200 import capi
    d1, d2 = elod_problems(name)
    o1, o2 = orcapi.Problem(d1), orcapi.Problem(d2)
    p1, p2 = capi.Plan(d1), capi.Plan(d2)
    R = 96
    log_sum, count, probs = capi.elod_run(p1, p2, R, seed=5, chain_id=2, want_probs=True)
    graphs = capi.debug_elod_graphs(p1, 0, R, seed=5, chain_id=2)
    late = capi.debug_elod_graphs(p1, 40, 8, seed=5, chain_id=2)           # a window that does not start at replicate 0
    assert (late == graphs[40:48]).all()
    want = []
    for r in range(R):
        dg, v = orcapi.elod_replicate(o1, o2, r, 5, 2)
        assert (graphs[r] == dg).all(), (name, r)
        want.append(v)
    want = np.array(want)
    scale = np.abs(want).max() + abs(o2.marker_transmission())
    assert np.abs(probs - want).max() <= 1e-12 * scale
    assert count == R
    ref_sum = want[0]
    for v in want[1:]:
        ref_sum = orcapi.log_sum(ref_sum, v)                               # LODscores::add, replicate by replicate
    assert abs(log_sum - ref_sum) <= 1e-12 * abs(ref_sum)
    # the trait likelihood that normalises the estimate
    assert abs(p2.trait_likelihood() - o2.trait_prob()) <= 1e-12 * abs(o2.trait_prob())
    p1.close(); p2.close()


@pytest.mark.skipif(not ref_available(), reason="example inputs live in oracle/_ref/examples")
@pytest.mark.parametrize("case", ["east", "loop", "xlinked", "east_dominant", "east_dominant_affected_only"])
def test_elod_agrees_with_reference_within_mc_error(case):
    """swiftlink::Elod (fake map, genotype-free pedigree, peel search, batched replicates) vs the reference's Elod::run"""
    from oracle import refapi
    from swiftlink_b200 import host as H
    ref = np.load(os.path.join(GOLDEN, "elod_ref.npz"))
    kw = dict(replicates=400000, peel_iterations=20000, seed=77)
    if case.startswith("east_dominant"):
        # `-a`: the simulation keeps the real phenotypes (slk_problem.person_prior), the scoring treats unaffected people as unknown
        total, per = H.elod(refapi.example("east")[0], frequency=1e-3, penetrance=(0.01, 0.8, 0.8), separation=0.1,
                            affected_only=case.endswith("affected_only"), **kw)
    else:
        total, per = H.elod(refapi.example(case)[0], sex_linked=(case == "xlinked"), **kw)
    vals = ref[case]
    # the reference values are means of 50 000 replicates each; ours of 400 000
    se = vals.std(ddof=1) * np.sqrt(1.0 / len(vals) + float(ref["replicates"]) / kw["replicates"])
    print("%s: device ELOD %.5f, reference %.5f +- %.5f (%d x %d replicates)" %
          (case, total, vals.mean(), vals.std(ddof=1), len(vals), int(ref["replicates"])))
    assert total == per[0]
    assert abs(total - vals.mean()) < 5.0 * se + 2e-4
