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
