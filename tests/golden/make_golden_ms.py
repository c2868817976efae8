"""Generates tests/golden/<case>_ms.npz: M-sampler / founder-allele-graph outputs of the UNMODIFIED
reference (oracle/_ref/libswiftref.so), on the descent graphs already frozen in <case>.npz.

    python tests/golden/make_golden_ms.py

Per case, for each frozen descent graph g:
  * fag_edges_g [M][2N], fag_lik_g [M]: FounderAlleleGraph4::reset + likelihood of every locus
    (founder_allele_graph4.cc:548-572, :34-424);
  * fag_flip_lik_g [M][K]: likelihood after FounderAlleleGraph4::flip of each of K sample meioses;
  * dg_likelihood is already in <case>.npz;
and for graph 2 a trace of MeiosisSampler::reset + step over every meiosis of m_ordering in pedigree
order then in reverse order (meiosis_sampler.cc:17-191): the uniforms the reference's mt19937 supplied,
raw_matrix and fb_matrix after every step, and the descent graph after every step.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import refapi as R          # noqa: E402
from common import golden, problem, case_files, FORCE_X, CASES   # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    R.set_threads(1)
    for name in CASES:
        R.seed(20261018)
        fx = golden(name)
        r = R.Ref(*case_files(name), sex_linked=bool(FORCE_X[name]))
        assert r.set_peel(np.array([o["peelnode"] for o in problem(name)["ops"]], np.uint32))
        out = {}
        order = r.ms_ordering()
        out["ms_ordering"] = order
        sample = order[:: max(1, len(order) // 6)]
        out["flip_meioses"] = sample
        for gi in range(len(fx["dgs"])):
            r.dg_set(fx["dgs"][gi])
            edges, lik, flik = [], [], []
            for l in range(r.M):
                e, v = r.fag(l)
                edges.append(e)
                lik.append(v)
                flik.append([r.fag(l, (r.F + m // 2, m % 2))[1] for m in sample])
            out["fag_edges_%d" % gi] = np.stack(edges).astype(np.int16)
            out["fag_lik_%d" % gi] = np.array(lik)
            out["fag_flip_lik_%d" % gi] = np.array(flik)
        r.dg_set(fx["dgs"][2])
        r.ms_reset(order[0])
        visit = list(order) + list(order[::-1])
        us, raws, fbs, dgs = [], [], [], []
        for m in visit:
            u, raw, fb = r.ms_step(m)
            us.append(u); raws.append(raw); fbs.append(fb); dgs.append(r.dg_get().astype(np.int8))
        out["trace_visit"] = np.array(visit, np.int32)
        out["trace_us"] = np.stack(us)
        out["trace_raw"] = np.stack(raws)
        out["trace_fb"] = np.stack(fbs)
        out["trace_dg"] = np.stack(dgs)
        path = os.path.join(OUT, name + "_ms.npz")
        np.savez_compressed(path, **out)
        print("%s -> %s (%.1f KB)" % (name, path, os.path.getsize(path) / 1024.0))
        r.close()


if __name__ == "__main__":
    main()
