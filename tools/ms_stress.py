"""M-sampler / L-sampler sweeps on the bench pedigree (200 members, first `--markers` SNPs) against the C oracle,
graph for graph (tuning / regression aid; the GPU tests do the same on the smaller golden pedigrees)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--markers", type=int, default=2000)
    ap.add_argument("--sweeps", type=int, default=6)
    ap.add_argument("--seed", type=int, default=77)
    args = ap.parse_args()
    from oracle import orcapi
    from swiftlink_b200 import capi, host as H
    paths = bench.workload_files(args.markers, "stress")
    hst = H.Host(*paths, lodscores=bench.N_LOD)
    assert hst.set_peel_by_names(bench.load_order()["order"])
    d = hst.problem_dict()
    orc = orcapi.Problem(d)
    plan = capi.Plan(d)
    ch = capi.Chain(plan, seed=args.seed, chain_id=3)
    ch.sequential_imputation(0, hst.M // 2)
    ref = ch.dg_download()
    bad = 0
    for it in range(1, 1 + args.sweeps):
        t0 = time.time()
        if it % 3 == 0:
            assert orc.ls_sweep(ref, args.seed, 3, it) == 0
            ch.lsampler_sweep(it)
            kind = "L"
        else:
            assert orc.ms_sweep(ref, args.seed, 3, it) == 0
            ch.msampler_sweep(it)
            kind = "M"
        got = ch.dg_download()
        diff = int((got != ref).sum())
        bad += diff
        print("sweep %d (%s): %d differing indicators, oracle+device %.1f s" % (it, kind, diff, time.time() - t0))
        if diff:
            ref = got.copy()          # keep going from the device's graph so later sweeps are still comparable
    want = orc.dg_likelihood(ref)
    got = ch.dg_likelihood()
    print("ln L(graph): device %.9f oracle %.9f rel %.2e" % (got, want, abs(got - want) / abs(want)))
    print("STRESS %s" % ("OK" if bad == 0 else "MISMATCH"))


if __name__ == "__main__":
    main()
