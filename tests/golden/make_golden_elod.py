"""Generates tests/golden/elod_ref.npz: the reference's own ELOD estimates (Elod::run, elod.cc:19-85) on its example
pedigrees, several seeds each, for the Monte-Carlo-error comparison in tests/test_gpu_elod.py.

    python tests/golden/make_golden_elod.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import refapi as R          # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REPLICATES, SEEDS = 50000, [11, 12, 13, 14, 15, 16]


def main():
    R.set_threads(1)
    R.seed(1)
    out = dict(replicates=REPLICATES, seeds=np.array(SEEDS))
    for name, x in (("east", False), ("loop", False), ("xlinked", True)):
        vals = [R.elod(R.example(name)[0], replicates=REPLICATES, sex_linked=x, seed=s) for s in SEEDS]
        out[name] = np.array(vals)
        print(name, vals)
    # a dominant model with reduced penetrance on east
    vals = [R.elod(R.example("east")[0], frequency=1e-3, penetrance=(0.01, 0.8, 0.8), separation=0.1, replicates=REPLICATES, seed=s)
            for s in SEEDS]
    out["east_dominant"] = np.array(vals)
    print("east_dominant", vals)
    # --elod -a on east with a reduced-penetrance model: unaffected people are simulated with their phenotype and scored
    # as unknown (elod.h:83-87)
    vals = [R.elod(R.example("east")[0], frequency=1e-3, penetrance=(0.01, 0.8, 0.8), separation=0.1, replicates=REPLICATES, seed=s,
                   affected_only=True) for s in SEEDS]
    out["east_dominant_affected_only"] = np.array(vals)
    print("east_dominant_affected_only", vals)
    np.savez_compressed(os.path.join(OUT, "elod_ref.npz"), **out)


if __name__ == "__main__":
    main()
