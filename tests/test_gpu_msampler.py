"""M-sampler and descent-graph likelihood on the device, through the C ABI, against the C oracle
(oracle/msampler_oracle.c, pinned bit-exact to the compiled reference) and the golden vectors.

Bar: founder-allele-graph edge lists bit-exact (integer labels); per-locus likelihoods within 1e-12 on
ln L (the device counts the exponents of the two allele frequencies exactly and rounds once, the
reference multiplies as it walks the graph); sampled descent graphs identical under the shared Philox
draws; the forward matrix within 1e-10 relative (blocked scan instead of a sequential pass; north_star
allows 1e-9); ln-likelihoods of whole graphs within 1e-12 relative."""
import numpy as np
import pytest

from common import CASES, golden, problem, oracle_problem

pytestmark = pytest.mark.gpu

TOL = 1e-12
TOL_FB = 1e-10


def ln_close(ln_got, lik_want):
    """device ln L against a reference likelihood (0 <-> -inf)"""
    ln_got, lik_want = np.asarray(ln_got, float), np.asarray(lik_want, float)
    zero = lik_want == 0.0
    if not (np.isneginf(ln_got) == zero).all():
        return False
    want = np.log(lik_want[~zero])
    return bool((np.abs(ln_got[~zero] - want) <= TOL * np.maximum(1.0, np.abs(want))).all())


@pytest.fixture(scope="module")
def gpu():
    from swiftlink_b200 import capi
    assert capi.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    made = {}

    def get(name, seed=77, chain_id=3):
        key = (name, seed, chain_id)
        if key not in made:
            plan = made.get(("plan", name))
            if plan is None:
                plan = capi.Plan(problem(name))
                made[("plan", name)] = plan
            made[key] = capi.Chain(plan, seed=seed, chain_id=chain_id)
        return made[key]
    yield get
    for k, v in made.items():
        if k[0] != "plan":
            v.close()
    for k, v in made.items():
        if k[0] == "plan":
            v.close()


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    d = np.abs(a - b)
    s = np.maximum(np.abs(a), np.abs(b))
    return float(np.max(np.where(s > 0, d / np.where(s > 0, s, 1), 0.0)))


@pytest.mark.parametrize("name", CASES)
def test_ordering(gpu, name):
    ch, ms = gpu(name), golden(name + "_ms")
    assert (ch.plan.msampler_ordering() == ms["ms_ordering"]).all()


@pytest.mark.parametrize("name", CASES)
def test_founder_allele_graph_bit_exact(gpu, name):
    """edge lists and likelihoods of every locus == reference (golden), plain and with a meiosis flipped"""
    fx, ms, ch = golden(name), golden(name + "_ms"), gpu(name)
    for gi in range(fx["dgs"].shape[0]):
        ch.dg_upload(fx["dgs"][gi])
        lnl, edges = ch.debug_fag(-1, edges=True)
        assert (edges == ms["fag_edges_%d" % gi]).all()
        assert ln_close(lnl, ms["fag_lik_%d" % gi])
        for k, m in enumerate(ms["flip_meioses"]):
            fl, _ = ch.debug_fag(int(m))
            assert ln_close(fl, ms["fag_flip_lik_%d" % gi][:, k]), (name, gi, m)
        assert (ch.dg_download() == fx["dgs"][gi]).all()


@pytest.mark.parametrize("name", CASES)
def test_flipped_edges_match_oracle(gpu, name):
    fx, orc, ch = golden(name), oracle_problem(name), gpu(name)
    dg = np.ascontiguousarray(fx["dgs"][3])
    ch.dg_upload(dg)
    for m in orc.ms_ordering():
        lnl, edges = ch.debug_fag(int(m), edges=True)
        want = []
        for l in range(orc.M):
            e, v = orc.fag(dg, l, (orc.F + int(m) // 2, int(m) % 2))
            assert (e == edges[l]).all()
            want.append(v)
        assert ln_close(lnl, want), (name, m)


@pytest.mark.parametrize("name", CASES)
def test_dg_likelihood(gpu, name):
    fx, ch = golden(name), gpu(name)
    for gi in range(fx["dgs"].shape[0]):
        ch.dg_upload(fx["dgs"][gi])
        got, want = ch.dg_likelihood(), float(fx["dg_likelihood"][gi])
        assert abs(got - want) <= TOL * abs(want)


@pytest.mark.parametrize("name", CASES)
def test_steps_match_oracle(gpu, name):
    """reset + one step per meiosis, forwards then backwards through the ordering: forward matrix,
    carried likelihoods and the sampled graph after every step"""
    fx, orc, ch = golden(name), oracle_problem(name), gpu(name, seed=4242, chain_id=1)
    for gi in (1, 2):
        dg = np.ascontiguousarray(fx["dgs"][gi]).copy()
        ch.dg_upload(dg)
        order = orc.ms_ordering()
        s = orc.msampler()
        assert s.reset(dg, int(order[0])) == 0
        ch.msampler_reset()
        flips = 0
        for it, m in enumerate(list(order) + list(order[::-1])):
            before = dg.copy()
            assert s.step(dg, int(m), 4242, 1, 100 + it) == 0
            ch.msampler_step(100 + it, int(m))
            fb, cur = ch.debug_msampler_state()
            st = s.state()
            assert rel(fb, st["fwd"]) <= TOL_FB, (name, gi, m, rel(fb, st["fwd"]))
            got = ch.dg_download()
            assert (got == dg).all(), (name, gi, it, m)
            person, par = orc.F + int(m) // 2, int(m) % 2
            assert ln_close(cur, st["raw"][np.arange(orc.M), dg[:, person, par]])
            flips += int((before != dg).sum())
        if name != "loop":
            assert flips > 0


@pytest.mark.parametrize("name", CASES)
def test_sweeps_match_oracle(gpu, name):
    """whole M-sweeps (shuffle, reset, every meiosis), interleaved with L-sweeps as MarkovChain::run
    mixes them: identical graphs"""
    fx, orc, ch = golden(name), oracle_problem(name), gpu(name, seed=99, chain_id=0)
    dg = np.ascontiguousarray(fx["dgs"][2]).copy()
    ch.dg_upload(dg)
    kinds = []
    for it in range(12):
        is_l = ch.sweep_is_lsampler(it, 0.5)
        kinds.append(is_l)
        if is_l:
            assert orc.ls_sweep(dg, 99, 0, it) == 0
            ch.lsampler_sweep(it)
        else:
            assert orc.ms_sweep(dg, 99, 0, it) == 0
            ch.msampler_sweep(it)
        assert (ch.dg_download() == dg).all(), (name, it, is_l)
    assert any(kinds) and not all(kinds)
    want = orc.dg_likelihood(dg)
    assert abs(ch.dg_likelihood() - want) <= TOL * abs(want)


def test_illegal_graph_is_reported(gpu):
    """a graph whose founder allele graph has likelihood 0 -> SLK_ERR_ILLEGAL_GRAPH, as the reference
    aborts with "illegal descent graph given to m-sampler" (meiosis_sampler.cc:31-34)"""
    from swiftlink_b200 import capi
    fx, orc, ch = golden("east"), oracle_problem("east"), gpu("east", seed=5, chain_id=9)
    dg = np.ascontiguousarray(fx["dgs"][2]).copy()
    # flip indicator bits until some locus becomes impossible (two flips: the first step only sees one of them)
    bad = []
    for m in orc.ms_ordering():
        person, par = orc.F + int(m) // 2, int(m) % 2
        for l in range(orc.M):
            if orc.fag(dg, l, (person, par))[1] == 0.0:
                bad.append((l, person, par))
        if len(bad) >= 2:
            break
    assert len(bad) >= 2
    for l, person, par in bad[:2]:
        dg[l, person, par] ^= 1
    assert orc.dg_sum_prior_prob(dg) < -1e300
    ch.dg_upload(dg)
    assert ch.dg_likelihood() == -np.finfo(float).max
    other = [int(m) for m in orc.ms_ordering() if (orc.F + int(m) // 2, int(m) % 2) not in [(b[1], b[2]) for b in bad[:2]]]
    with pytest.raises(capi.SlkError) as e:
        ch.msampler_step(0, other[0])
        ch.sync()
    assert e.value.code == -7
    ch.dg_upload(fx["dgs"][2])
    ch.sync()


def test_dg_swap(gpu):
    """slk_dg_swap exchanges two chains' graphs in O(1) and invalidates the carried likelihoods"""
    fx = golden("east")
    a, b = gpu("east", seed=1, chain_id=0), gpu("east", seed=1, chain_id=1)
    a.dg_upload(fx["dgs"][1]); b.dg_upload(fx["dgs"][3])
    la, lb = a.dg_likelihood(), b.dg_likelihood()
    a.dg_swap(b)
    assert (a.dg_download() == fx["dgs"][3]).all() and (b.dg_download() == fx["dgs"][1]).all()
    assert a.dg_likelihood() == lb and b.dg_likelihood() == la
    a.dg_swap(b)
    assert (a.dg_download() == fx["dgs"][1]).all()


@pytest.mark.parametrize("name", ["east", "xlinked"])
def test_mc3_ladder_runs_and_scores(name, tmp_path):
    """Mc3::run on the device (three heated chains, default sampler mix): swap bookkeeping and a cold-chain
    LOD curve in the band of the reference's own default-mix replicates"""
    from common import case_files, FORCE_X, ref_available
    if not ref_available():
        pytest.skip("example inputs live in oracle/_ref/examples")
    from swiftlink_b200 import host as H
    fx = golden(name)
    h = H.Host(*case_files(name, tmp_path), sex_linked=bool(FORCE_X[name]))
    assert h.set_peel([o["peelnode"] for o in problem(name)["ops"]])
    out = h.run_mc3(3, burnin=300, iterations=900, exchange_period=10, seed=31, si_iterations=20)
    tries = out["swap_success"] + out["swap_failure"]
    assert tries.sum() == 120 and (tries > 0).all()
    assert out["swap_success"].sum() > 0                      # temperatures 0.998 / 0.996: swaps are nearly free
    lod = out["lod"]
    assert np.isfinite(lod).all()
    ref = fx["lod_curves_default_mix"].mean(axis=(1, 2))
    assert abs(lod.mean() - ref.mean()) < 6.0 * ref.std(ddof=1) + 0.75
    h.close()


def test_replicates_in_flight_do_not_change_the_result(tmp_path):
    """LinkageProgram::run_pedigree's -R loop (linkage_program.cc:96-108): three replicate chains advanced in turn on
    three streams give bit for bit the merged LOD table of the same replicates run one after the other (a chain's draws
    are keyed by seed, chain id and iteration, not by when its kernels run)"""
    from common import case_files, FORCE_X, ref_available
    if not ref_available():
        pytest.skip("example inputs live in oracle/_ref/examples")
    from swiftlink_b200 import host as H
    h = H.Host(*case_files("east", tmp_path), sex_linked=bool(FORCE_X["east"]))
    assert h.set_peel([o["peelnode"] for o in problem("east")["ops"]])
    one = h.run_replicates(3, 1, burnin=60, iterations=240, seed=77)
    many = h.run_replicates(3, 3, burnin=60, iterations=240, seed=77)
    two = h.run_replicates(3, 2, burnin=60, iterations=240, seed=77)
    assert np.isfinite(one).all()
    assert (one == many).all() and (one == two).all()
    h.close()


def test_bench_pedigree_sweeps_match_oracle():
    """the bench workload's pedigree (200 members, 53 founders, 107 typed, cutsets up to 6; first 1 500 SNPs): M-sweeps
    (939-CTA-scale likelihood launches overlapped with the cluster chain kernel) and L-sweeps reproduce the oracle's
    graphs bit for bit, and the graph likelihood agrees"""
    import sys
    from common import ROOT
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench
    from oracle import orcapi
    from swiftlink_b200 import capi, host as H
    hst = H.Host(*bench.workload_files(1500, "t_sweeps"), lodscores=bench.N_LOD)
    assert hst.set_peel_by_names(bench.load_order()["order"])
    d = hst.problem_dict()
    orc, plan = orcapi.Problem(d), capi.Plan(d)
    ch = capi.Chain(plan, seed=77, chain_id=3)
    ch.sequential_imputation(0, hst.M // 2)
    ref = ch.dg_download()
    for it in (1, 2, 3, 4):
        if it == 3:
            assert orc.ls_sweep(ref, 77, 3, it) == 0
            ch.lsampler_sweep(it)
        else:
            assert orc.ms_sweep(ref, 77, 3, it) == 0
            ch.msampler_sweep(it)
        assert (ch.dg_download() == ref).all(), it
    want = orc.dg_likelihood(ref)
    assert abs(ch.dg_likelihood() - want) <= TOL * abs(want)
    ch.close(); plan.close(); hst.close()


@pytest.mark.parametrize("env", ["SLK_MS_NO_REC", "SLK_NO_PDL", "SLK_MS_NO_PREFIX", "SLK_MS_RUN_AHEAD", "SLK_MS_SNAPSHOT=0", "SLK_MS_SNAPSHOT=2"])
def test_msampler_fallback_paths(env):
    """the library's alternative launch paths -- launch record derived inside the kernel instead of on the host, plain
    launches instead of programmatic dependent ones, nothing running ahead of the predecessor, one likelihood launch
    in flight instead of two, no forest snapshots / a snapshot every second launch -- give the same graphs
    (the switches are read once per process, hence the subprocess)"""
    import os
    import subprocess
    import sys
    from common import ROOT
    e = dict(os.environ)
    e[env.split("=")[0]] = env.split("=")[1] if "=" in env else "1"
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_msampler.py"), "-q", "-x",
                          "-k", "test_sweeps_match_oracle or test_steps_match_oracle"], env=e, capture_output=True, text=True,
                         timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "passed" in out.stdout


def _bench_problem(markers, tag):
    import sys
    from common import ROOT
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench
    from oracle import orcapi
    from swiftlink_b200 import capi, host as H
    hst = H.Host(*bench.workload_files(markers, tag), lodscores=bench.N_LOD)
    assert hst.set_peel_by_names(bench.load_order()["order"])
    d = hst.problem_dict()
    return hst, orcapi.Problem(d), capi.Plan(d)


def _check_lod_sample(ch, orc, dg, intervals):
    """the production scoring kernel (slk_lodscore_kernel<T, CTA, false>, the instantiation bench.py times) over ALL
    positions, compared with the oracle's Peeler::process at a sample of intervals"""
    ch.lodscore_init()
    ch.lodscore_accumulate()
    raw, count = ch.lodscore_read()
    assert count == 1
    for l in intervals:
        prob = orc.lod_interval(dg, int(l))[1]
        assert np.abs(raw[l] - prob).max() <= 1e-12 * np.abs(prob).max(), l
    ch.lodscore_accumulate()                           # LODscores::add: log_sum of the same value twice = value + ln 2
    raw2, count = ch.lodscore_read()
    assert count == 2
    assert np.abs(raw2 - (raw + np.log(2.0))).max() <= 1e-12 * np.abs(raw).max()


def test_bench_pedigree_lod_pass_matches_oracle():
    """bench pedigree, first 1 500 SNPs: a scoring pass of the hot (non-debug) kernel at the production team geometry
    against the oracle (1e-12 relative), after L-sweeps of the hot sampler kernel that the oracle reproduces bit for bit"""
    hst, orc, plan = _bench_problem(1500, "t_lod")
    ch = capi_chain(plan, 77, 3)
    ch.sequential_imputation(0, hst.M // 2)
    ref = ch.dg_download()
    for it in (1, 2, 3):                              # several L-sweeps at the production geometry
        assert orc.ls_sweep(ref, 77, 3, it) == 0
        ch.lsampler_sweep(it)
        assert (ch.dg_download() == ref).all(), it
    _check_lod_sample(ch, orc, ref, np.linspace(0, hst.M - 2, 40).astype(int))
    ch.close(); plan.close(); hst.close()


@pytest.mark.slow
def test_bench_workload_full_size_matches_oracle():
    """the bench workload itself (200 members x 10 000 SNPs): one M-sweep, two L-sweeps and a scoring pass of the
    production kernels against the oracle -- graphs bit for bit, ln L(graph) and sampled LOD terms at 1e-12"""
    hst, orc, plan = _bench_problem(10000, "t_full")
    ch = capi_chain(plan, 20261017, 0)
    ch.sequential_imputation(0, hst.M // 2)
    ref = ch.dg_download()
    for it, kind in ((1, "M"), (2, "L"), (3, "L")):
        if kind == "L":
            assert orc.ls_sweep(ref, 20261017, 0, it) == 0
            ch.lsampler_sweep(it)
        else:
            assert orc.ms_sweep(ref, 20261017, 0, it) == 0
            ch.msampler_sweep(it)
        assert (ch.dg_download() == ref).all(), (it, kind)
    want = orc.dg_likelihood(ref)
    assert abs(ch.dg_likelihood() - want) <= TOL * abs(want)
    _check_lod_sample(ch, orc, ref, np.linspace(0, hst.M - 2, 24).astype(int))
    ch.close(); plan.close(); hst.close()


def capi_chain(plan, seed, chain_id):
    from swiftlink_b200 import capi
    return capi.Chain(plan, seed=seed, chain_id=chain_id)
