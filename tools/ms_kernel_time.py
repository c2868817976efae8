"""Standalone duration of the M-sampler step / chain kernels on the bench workload (tuning aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from swiftlink_b200 import capi, host as H
paths = bench.workload_files(bench.N_MARKERS, "mk")
hst = H.Host(*paths, lodscores=bench.N_LOD)
assert hst.set_peel_by_names(bench.load_order()["order"])
plan = H.PlanFromHost(hst, device=0)
chain = capi.Chain(plan, seed=20261017, chain_id=0)
chain.sequential_imputation(run=0, start_locus=hst.M // 2)
for it in range(1, 6):
    chain.lsampler_sweep(it)
chain.msampler_sweep(7)
chain.sync()
order = plan.msampler_ordering()
for which, name in ((0, "step"), (1, "chain")):
    chain.debug_msampler_launch(int(order[0]), int(order[1]), which, 5); chain.sync()
    t = time.time()
    chain.debug_msampler_launch(int(order[0]), int(order[1]), which, 200); chain.sync()
    print("%s kernel: %.2f us per launch (200 back to back)" % (name, 1e6 * (time.time() - t) / 200))
import numpy as np
tr = chain.debug_msampler_trace(int(order[2]), int(order[3]))
for row in tr[:4]:
    if row[0] > 0 and row[6] > 0:
        print("   step kernel cycles [prologue, walk, finish]", [int(row[4] - row[0]), int(row[5] - row[4]), int(row[6] - row[5])])
