"""The multi-GPU product path: swiftlink::ReplicateJob (one device's share of a -R job, plain chains or MC3 ladders, all
resident at once) and `python -m swiftlink_b200.run` (replicates dealt out over the ranks, tables and swap counters
merged by all-reduces; LinkageProgram::run_pedigree's loop, linkage_program.cc:96-108, Mc3::run, mc3.cc:81-200,
LODscores::merge_results, lod_score.h:98-105)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from common import ROOT, FORCE_X, case_files, problem, ref_available

pytestmark = pytest.mark.gpu


def _host(name, tmp_path):
    from swiftlink_b200 import host as H
    h = H.Host(*case_files(name, tmp_path), sex_linked=bool(FORCE_X[name]))
    assert h.set_peel([o["peelnode"] for o in problem(name)["ops"]])
    return h


def _normalised(r):
    from swiftlink_b200 import dist as sdist
    import torch
    return sdist.normalise(torch.from_numpy(r["raw"]), r["count"], r["trait_prob"]).numpy()


def test_job_of_plain_replicates_matches_the_replicate_loop(tmp_path):
    """a job of replicates {0, 1, 2} advanced in two pieces = swiftlink::run_replicates(3), bit for bit"""
    if not ref_available():
        pytest.skip("example inputs live in oracle/_ref/examples")
    from swiftlink_b200 import host as H
    h = _host("east", tmp_path)
    want = h.run_replicates(3, 3, burnin=40, iterations=200, seed=77, si_iterations=10)
    job = H.Job(h, [0, 1, 2], 40, 200, seed=77, si_iterations=10)
    assert job.advance(100) == 100 and job.advance(1000) == 140 and job.advance(10) == 0
    r = job.results()
    job.close()
    assert r["count"] == 3 * 20
    assert (_normalised(r).reshape(want.shape) == want).all()
    h.close()


def test_job_of_ladders_matches_mc3_run(tmp_path):
    """a job of two MC3 ladders in flight = two Mc3::run calls one after the other: cold-chain tables merged, swap
    counters summed"""
    if not ref_available():
        pytest.skip("example inputs live in oracle/_ref/examples")
    from swiftlink_b200 import host as H
    import torch
    from swiftlink_b200 import dist as sdist
    h = _host("east", tmp_path)
    kw = dict(burnin=40, iterations=160, exchange_period=10, seed=31, si_iterations=5)
    a = h.run_mc3(3, chain_id=4, **kw)
    b = h.run_mc3(3, chain_id=7, **kw)
    job = H.Job(h, [4, 7], 40, 160, seed=31, si_iterations=5, mc3_chains=3, exchange_period=10)
    assert job.advance(95) == 90                      # whole spurts of the exchange period
    assert job.advance(1000) == 110
    r = job.results()
    job.close()
    assert (r["swap_success"][:2] == a["swap_success"] + b["swap_success"]).all()
    assert (r["swap_failure"][:2] == a["swap_failure"] + b["swap_failure"]).all()
    assert int((r["swap_success"] + r["swap_failure"])[:2].sum()) == 2 * 20
    # merge of the two normalised tables: log10-mean of the likelihood ratios
    want = np.log10((10.0 ** a["lod"] + 10.0 ** b["lod"]) / 2.0)
    got = _normalised(r).reshape(want.shape)
    assert np.abs(got - want).max() <= 1e-9 * max(1.0, np.abs(want).max())
    h.close()


def test_two_ranks_give_the_single_process_table(tmp_path):
    """`python -m swiftlink_b200.run` with two processes (replicates 0, 2 on rank 0 and 1, 3 on rank 1; gloo all-reduces,
    both ranks on this box's GPU) writes the table a single process writes"""
    if not ref_available():
        pytest.skip("example inputs live in oracle/_ref/examples")
    ped, mp, dat = case_files("east", tmp_path)
    common = ["-p", ped, "-m", mp, "-d", dat, "-R", "4", "-b", "30", "-i", "120", "-s", "5", "-q", "20000", "-S", "5",
              "-M", "-z", "2", "-y", "10"]
    env = dict(os.environ, PYTHONPATH=ROOT)
    one = str(tmp_path / "one.out")
    two = str(tmp_path / "two.out")
    r1 = subprocess.run([sys.executable, "-m", "swiftlink_b200.run"] + common + ["-o", one], cwd=ROOT, env=env,
                        capture_output=True, text=True, timeout=600)
    assert r1.returncode == 0, r1.stdout[-1500:] + r1.stderr[-1500:]
    r2 = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                         "127.0.0.1", "--master-port", "29631", "-m", "swiftlink_b200.run"] + common +
                        ["-o", two, "--backend", "gloo"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r2.returncode == 0, r2.stdout[-1500:] + r2.stderr[-1500:]

    def table(fn):
        # linkage_writer.cc:14-92: "-<tab>position<tab>lod" per scored position
        return np.array([[float(x) for x in ln.split()[1:3]] for ln in open(fn) if ln.startswith("-")])
    a, b = table(one), table(two)
    assert a.shape == b.shape and a.size > 0
    assert np.abs(a - b).max() <= 1e-5 * max(1.0, np.abs(a).max())     # the file prints six significant digits
    assert "0 -- 1" in r1.stdout and "0 -- 1" in r2.stdout
