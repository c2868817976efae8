"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libswiftref.so).

Run in the build container (where /root/reference exists and `make -C oracle` has been run):

    python tests/golden/make_golden.py

Each fixture freezes, for one of the reference's example pedigrees:
  * every input table the hot path reads (pedigree, genotypes, map, disease model),
  * the elimination masks, the peel sequence and the derived PeelOperation fields,
  * the per-locus index tables of a few ops (bit-exact targets for validity derivation),
  * descent graphs taken from the reference's own chain (after sequential imputation and a
    short MCMC run),
  * for a sample of (graph, locus): all peel and presum matrices of the L-sampler forward pass,
  * for a sample of (graph, interval): result / prob of Peeler::process at every position and
    the trait peel matrices of one position,
  * ln P(T), and the final LOD curves of independent reference chains (seeds listed) for
    Monte-Carlo-error bands.
The reference has no tests or golden vectors of its own; these outputs of
the compiled reference are what pins oracle/peel_oracle.c and, through it, the CUDA path.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import refapi as R          # noqa: E402
from oracle import orcapi as O          # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
MAXC, MAXP, MAXK = 10, 8, 10

CASES = [
    # name, force -X, peel search iterations, sample loci, sample intervals, chain seeds
    ("loop", 0, 100000, [0, 1, 2], [0, 1], list(range(101, 111))),
    ("xlinked", 1, 100000, [0, 3, 9], [0, 4, 8], list(range(201, 211))),
    ("east", 0, 1000000, [0, 1, 37, 98, 99], [0, 50, 98], list(range(301, 311))),
    # not one of the reference's examples: a generated 64-member pedigree with 8 marriage loops,
    # cutsets up to 5, allele frequencies != 0.5 and untyped non-founders (swiftlink_b200/synth.py);
    # it is what exposes the reference's "everyone is a founder" marker prior (person.cc:224-299
    # called from pedigree_parser.cc:153) and exercises 1024-cell peel matrices
    ("inbred", 0, 200000, [0, 7, 15], [0, 8, 14], list(range(401, 406))),
]

sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import INBRED               # noqa: E402


def case_files(name):
    if name != "inbred":
        return R.example(name)
    import tempfile
    from swiftlink_b200 import synth
    ped = synth.generate(**INBRED)
    return synth.write_linkage(ped, os.path.join(tempfile.mkdtemp(prefix="slk_golden_"), "inbred"))


def pack_ops(ops):
    n = len(ops)
    out = dict(op_type=np.zeros(n, np.int32), op_peelnode=np.zeros(n, np.int32),
               op_ncut=np.zeros(n, np.int32), op_cutset=-np.ones((n, MAXC), np.int32),
               op_nprev=np.zeros(n, np.int32), op_prev=-np.ones((n, MAXP), np.int32),
               op_nchild=np.zeros(n, np.int32), op_children=-np.ones((n, MAXK), np.int32))
    for i, o in enumerate(ops):
        out["op_type"][i] = o["type"]
        out["op_peelnode"][i] = o["peelnode"]
        for key, cnt, arr in (("cutset", "op_ncut", "op_cutset"), ("previous", "op_nprev", "op_prev"),
                              ("children", "op_nchild", "op_children")):
            out[cnt][i] = len(o[key])
            out[arr][i, :len(o[key])] = o[key]
    return out


def main():
    only = sys.argv[1:]
    R.set_threads(1)
    for name, force_x, peel_iters, loci, intervals, seeds in CASES:
        R.seed(20261017)
        if only and name not in only:
            continue
        r = R.Ref(*case_files(name), sex_linked=bool(force_x))
        r.build_peel(peel_iters)
        d = O.problem_from_ref(r)
        fx = dict(N=d["N"], F=d["F"], M=d["M"], nlod=d["nlod"], sex_linked=d["sex_linked"])
        for k in ("mother", "father", "sex", "affection", "typed", "disease_prob", "marker_prob", "genotypes",
                  "elim", "theta", "partial", "gdist", "minor", "mapprob", "mapxprob"):
            fx[k] = d[k]
        fx.update(pack_ops(d["ops"]))
        fx["dm_freq"] = r.disease_model()["freq"]
        fx["dm_penetrance"] = r.disease_model()["penetrance"]
        fx["peel_cost"] = r.peel_cost()
        fx["person_names"] = np.array(r.person_names())
        fx["marker_names"] = np.array(r.marker_names())

        # index tables of every op at the sampled loci
        for i in range(r.num_ops()):
            fx["lod_indices_%d" % i] = r.op_indices(i, 0)
            for l in loci:
                fx["matrix_indices_%d_%d" % (i, l)] = r.op_indices(i, 1, l)
                fx["presum_indices_%d_%d" % (i, l)] = r.op_indices(i, 2, l)

        # descent graphs from the reference's own chain
        dgs = []
        r.dg_random()
        dgs.append(r.dg_get())
        r.sequential_imputation(50)
        dgs.append(r.dg_get())
        r.chain_run(200, 0)
        dgs.append(r.dg_get())
        r.chain_run(200, 0)
        dgs.append(r.dg_get())
        fx["dgs"] = np.stack(dgs)
        fx["dg_likelihood"] = np.array([(r.dg_set(g), r.dg_likelihood())[1] for g in dgs])

        fx["sample_loci"] = np.array(loci, np.int32)
        fx["sample_intervals"] = np.array(intervals, np.int32)
        for gi, g in enumerate(dgs):
            r.dg_set(g)
            for l in loci:
                res, mat, pre = r.ls_forward(l)
                fx["ls_result_%d_%d" % (gi, l)] = res
                fx["ls_mat_%d_%d" % (gi, l)] = mat
                fx["ls_pre_%d_%d" % (gi, l)] = pre
            # sequential-imputation mode of the same routine (sampler_rfunction.h:84-112)
            res, mat, pre = r.ls_forward(loci[1], mode=1, ignore_left=True, ignore_right=False)
            fx["ls_si_mat_%d" % gi] = mat
            for itv in intervals:
                res, prob, mat = r.lod_interval(itv, 2)
                fx["lod_result_%d_%d" % (gi, itv)] = res
                fx["lod_prob_%d_%d" % (gi, itv)] = prob
                fx["lod_mat_%d_%d" % (gi, itv)] = mat
            fx["recomb_%d" % gi] = np.array([r.dg_recombination_prob(l) for l in range(r.M - 1)])
            # one full scoring pass
            fx["lod_pass_%d" % gi] = np.stack([r.lod_interval(itv)[1] for itv in range(r.M - 1)])
        fx["marker_transmission"] = r.dg_marker_transmission()
        fx["trait_prob"] = r.calc_trait_prob()

        # final LOD curves of independent reference chains (L-sampler only and default mix)
        curves_l, curves_mix = [], []
        for s in seeds:
            R.seed(s)
            r.dg_random()
            r.sequential_imputation(100)
            curves_l.append(r.chain_run(2000, 4000, 10, 1.0)["lod"])
            R.seed(s + 5000)
            r.dg_random()
            r.sequential_imputation(100)
            curves_mix.append(r.chain_run(2000, 4000, 10, 0.5)["lod"])
        fx["chain_seeds"] = np.array(seeds)
        fx["lod_curves_lsampler_only"] = np.stack(curves_l)
        fx["lod_curves_default_mix"] = np.stack(curves_mix)

        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **fx)
        print("%s: N=%d M=%d cost=%d -> %s (%.1f KB)" % (name, r.N, r.M, r.peel_cost(), path,
                                                         os.path.getsize(path) / 1024.0))
        r.close()


if __name__ == "__main__":
    main()
