"""Shared helpers for the test-suite: golden fixtures -> problem dicts, oracle / reference access."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["loop", "xlinked", "east", "inbred"]
FORCE_X = {"loop": 0, "xlinked": 1, "east": 0, "inbred": 0}

# the generated pedigree behind tests/golden/inbred.npz (must match tests/golden/make_golden.py)
INBRED = dict(n_members=64, n_markers=16, seed=7, spacing_cm=0.8, n_generations=5, loops=(3, 8), min_generation=4,
              founder_frac=(0.2, 0.45), min_affected=1, cousin_prob=0.25)

_cache = {}


def golden(name):
    if name not in _cache:
        _cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    return _cache[name]


def unpack_ops(fx):
    ops = []
    for i in range(len(fx["op_type"])):
        ops.append(dict(type=int(fx["op_type"][i]), peelnode=int(fx["op_peelnode"][i]),
                        cutset=[int(x) for x in fx["op_cutset"][i, :fx["op_ncut"][i]]],
                        previous=[int(x) for x in fx["op_prev"][i, :fx["op_nprev"][i]]],
                        children=[int(x) for x in fx["op_children"][i, :fx["op_nchild"][i]]]))
    return ops


def problem(name):
    """problem dict in the layout of oracle.orcapi.problem_from_ref, from a golden fixture"""
    fx = golden(name)
    d = dict(N=int(fx["N"]), F=int(fx["F"]), M=int(fx["M"]), nlod=int(fx["nlod"]), sex_linked=int(fx["sex_linked"]))
    for k in ("mother", "father", "sex", "affection", "typed", "disease_prob", "marker_prob", "genotypes", "elim",
              "theta", "partial", "gdist", "minor", "mapprob", "mapxprob"):
        d[k] = fx[k]
    d["ops"] = unpack_ops(fx)
    return d


def oracle_problem(name):
    from oracle import orcapi
    key = ("orc", name)
    if key not in _cache:
        _cache[key] = orcapi.Problem(problem(name))
    return _cache[key]


def ref_available():
    from oracle import refapi
    return refapi.available() and os.path.isdir(refapi.EXAMPLES)


def matrix_offsets(ops):
    offs, o = [], 0
    for op in ops:
        offs.append(o)
        o += 4 ** len(op["cutset"])
    return offs


def case_files(name, tmpdir=None):
    """LINKAGE ped/map/dat of a test case: the reference's examples (shipped inside oracle/_ref)
    or, for `inbred`, files written by the generator"""
    if name == "inbred":
        import tempfile
        from swiftlink_b200 import synth
        d = tmpdir or tempfile.mkdtemp(prefix="slk_inbred_")
        return synth.write_linkage(synth.generate(**INBRED), os.path.join(str(d), "inbred"))
    from oracle import refapi
    return refapi.example(name)


def haldane(m):
    return 0.5 * (1.0 - np.exp(-2.0 * m))


def elod_problems(name, separation=0.05, trait_freq=1e-4):
    """The two problems of the ELOD calculation (elod.h:40-90) on a golden pedigree: the three-locus sampler
    problem (marker, simulated trait locus whose genotype prior is each person's disease probability, marker;
    no genotypes; priors with the resolved founder flags) and the two-marker trait problem.  Returns
    (device dict 1, device dict 2, oracle dict 1, oracle dict 2)."""
    from oracle import orcapi
    base = problem(name)
    N, F = base["N"], base["F"]
    assert not base["sex_linked"], "autosomal cases only (elimination masks are trivially 'everything legal')"

    def snp_prob(minor):
        major = 1.0 - minor
        p = np.array([major * major, minor * minor, minor * major, minor * major])     # UU, AA, AU, UA (genetic_map.h:68-74)
        return p / p.sum()

    th_half, th_full = haldane(separation / 2), haldane(separation)

    def partial(theta, n):
        return haldane((-0.5 * np.log(1 - 2 * theta)) / (n + 1))

    founder = (np.asarray(base["mother"]) < 0).astype(np.int32)
    out = []
    for M, minors, thetas in ((3, [0.5, trait_freq, 0.5], [th_half, th_half]), (2, [0.5, 0.5], [th_full])):
        d = dict(base)
        d.update(M=M, nlod=1, genotypes=np.zeros((N, M), np.int32), typed=np.zeros(N, np.int32),
                 elim=np.full((M, N), 15, np.int32), theta=np.array(thetas), partial=np.array([partial(t, 1) for t in thetas]),
                 minor=np.array(minors), mapprob=np.stack([snp_prob(m) for m in minors]),
                 mapxprob=np.stack([snp_prob(m) for m in minors]), prior_as_founder=founder)
        mp = np.zeros((N, M, 4))
        for i in range(N):
            for l in range(M):
                mp[i, l] = orcapi.marker_prob(int(founder[i]), 0, 0, 0, d["mapprob"][l])
        if M == 3:
            mp[:, 1, :] = np.asarray(base["disease_prob"])          # Person::copy_disease_probs(1)
            d["disease_prior_locus"] = 1
        d["marker_prob"] = mp
        out.append(d)
    return out[0], out[1]
