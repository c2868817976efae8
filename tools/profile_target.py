"""Small driver for ncu / timing experiments on the bench workload (not part of the product).

    python tools/profile_target.py [--markers M] [--sweeps S] [--lod L]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--markers", type=int, default=bench.N_MARKERS)
    ap.add_argument("--sweeps", type=int, default=5)
    ap.add_argument("--lod", type=int, default=1)
    ap.add_argument("--msweeps", type=int, default=0)
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--trace", action="store_true")
    ap.add_argument("--timeline", action="store_true", help="%%globaltimer stamps of the launches of one M-sweep")
    args = ap.parse_args()
    from swiftlink_b200 import capi, host as H
    paths = bench.workload_files(args.markers, "prof")
    hst = H.Host(*paths, lodscores=bench.N_LOD)
    assert hst.set_peel_by_names(bench.load_order()["order"])
    plan = H.PlanFromHost(hst)
    print({k: v for k, v in plan.stats().items()})
    chain = capi.Chain(plan, seed=1)
    chain.sequential_imputation(0, hst.M // 2)
    chain.sync()
    t0 = time.time()
    for it in range(args.sweeps):
        chain.lsampler_sweep(1 + it)
    chain.sync()
    t1 = time.time()
    for _ in range(args.lod):
        chain.lodscore_accumulate()
    chain.sync()
    t2 = time.time()
    if args.msweeps:
        chain.msampler_sweep(1000); chain.sync()
        t3 = time.time()
        for it in range(args.msweeps):
            chain.msampler_sweep(2000 + it)
        t_enq = time.time()
        chain.sync()
        t4 = time.time()
        n = len(plan.msampler_ordering())
        print("M-sweep ms %.3f (%d meioses, %.2f us per step; host enqueue %.3f ms per sweep); ln L = %.3f" %
              (1e3 * (t4 - t3) / args.msweeps, n, 1e6 * (t4 - t3) / args.msweeps / n, 1e3 * (t_enq - t3) / args.msweeps,
               chain.dg_likelihood()))
    if args.timeline:
        import numpy as np
        chain.msampler_sweep(900); chain.sync()
        tl, cta = chain.debug_msampler_timeline(901, cta_pair=100)
        t0 = float(tl[0, 0])
        us = lambda x: (float(x) - t0) * 1e-3
        for j in range(40, 56, 2):
            a, b = tl[j], tl[j + 1]
            print("pair %3d  step: start %7.1f walk %7.1f wait %7.1f..%7.1f end %7.1f | chain: start %7.1f wait %7.1f..%7.1f end %7.1f (us)"
                  % (j // 2, us(a[0]), us(a[1]), us(a[2]), us(a[3]), us(a[4]), us(b[0]), us(b[1]), us(b[2]), us(b[3])))
        live = cta[:, 1] != 0
        if live.any():
            first = float(cta[live, 0].min())
            st = np.sort((cta[live, 0].astype(np.float64) - first) * 1e-3)
            en = np.sort(((cta[live, 1] >> np.uint64(10)).astype(np.float64) - first) * 1e-3)
            q = lambda v, f: v[min(len(v) - 1, int(len(v) * f))]
            print("likelihood CTAs of pair 50 (%d): start p50 %.1f p99 %.1f max %.1f | end p10 %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f (us)"
                  % (len(st), q(st, .5), q(st, .99), st[-1], q(en, .1), q(en, .5), q(en, .9), q(en, .99), en[-1]))
    if args.msweeps and args.trace:
        import numpy as np
        order = plan.msampler_ordering()
        tr = chain.debug_msampler_trace(int(order[0]), int(order[1]))
        tr = chain.debug_msampler_trace(int(order[2]), int(order[3]))
        d = np.diff(tr[:12, :7], axis=1)
        for nm, base in (("CTA 0", 96), ("last CTA", 120)):
            st = tr.ravel()[base:base + 17]
            print("chain kernel %s cycles [stage, product, cta scan, cluster sync, recurrence, map scans, apply, sync] x 2 steps: %s  total %d" %
                  (nm, np.diff(st).tolist(), int(st[-1] - st[0])))
        print("step kernel cycles [prologue (tables, slot mask), walk, finish + write-back] for sampled warps:")
        for row in tr[:12]:
            if row[0] > 0 and row[6] > 0:
                print("   ", [int(row[4] - row[0]), int(row[5] - row[4]), int(row[6] - row[5])], "total", int(row[6] - row[0]))
    if args.trace and args.sweeps:
        import numpy as np
        st = plan.stats()
        t = chain.debug_trace(99, 0)
        d = np.diff(t)
        nf, nb = int(st["ls_flevels"]), int(st["ls_blevels"])
        print("stage %d | forward levels (sum %d): %s | backward levels (sum %d): %s | indicators %d | total %d cycles" %
              (d[0], d[1:1 + nf].sum(), d[1:1 + nf].tolist(), d[1 + nf:1 + nf + nb].sum(), d[1 + nf:1 + nf + nb].tolist(),
               d[1 + nf + nb], t[-1] - t[0]))
    if args.time:
        print("sweep ms %.3f   lod pass ms %.3f" % (1e3 * (t1 - t0) / max(args.sweeps, 1), 1e3 * (t2 - t1) / max(args.lod, 1)))


if __name__ == "__main__":
    main()
