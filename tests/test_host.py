"""The C++ host side (parsers, pedigree/map/disease-model tables, genotype elimination,
peel-sequence generator) against the golden fixtures frozen from the reference and, when
oracle/_ref is present, against the live reference on the same files."""
import os

import numpy as np
import pytest

from common import CASES, FORCE_X, case_files, golden, problem, unpack_ops, ref_available
from swiftlink_b200 import capi, host as H

needs_ref = pytest.mark.skipif(not ref_available(), reason="oracle/_ref (compiled reference + example inputs) not present")

X = FORCE_X
example_files = case_files


def test_host_symbols_exported():
    import ctypes
    import re
    from common import ROOT
    L = ctypes.CDLL(capi.LIB_PATH)
    text = open(os.path.join(ROOT, "include", "swiftlink_b200_host.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = sorted(set(re.findall(r"\b(slk_host_[a-z0-9_]+)\s*\(", text)))
    assert declared == sorted(H.SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name


@needs_ref
@pytest.mark.parametrize("name", CASES)
def test_tables_and_plan_match_golden(name):
    """everything the hot path reads, parsed by the host code from the example files, is
    bit-identical to what the reference produced (golden fixture)"""
    fx = golden(name)
    h = H.Host(*example_files(name), sex_linked=bool(X[name]))
    assert (h.N, h.F, h.M, h.nlod, h.sex_linked) == tuple(int(fx[k]) for k in ("N", "F", "M", "nlod", "sex_linked"))
    assert h.person_names() == [str(s) for s in fx["person_names"]]
    assert h.marker_names() == [str(s) for s in fx["marker_names"]]
    pt = h.person_table()
    for k in ("mother", "father", "sex", "affection", "typed", "disease_prob"):
        assert (pt[k] == fx[k]).all(), k
    assert (h.genotypes() == fx["genotypes"]).all()
    assert (h.marker_trait_prob() == fx["marker_prob"]).all()
    mt = h.map_table()
    for a, b in (("gdist", "gdist"), ("minor", "minor"), ("prob", "mapprob"), ("xprob", "mapxprob"),
                 ("theta", "theta"), ("partial", "partial")):
        assert (mt[a] == fx[b]).all(), a
    dm = h.disease_model()
    assert dm["freq"] == float(fx["dm_freq"]) and (dm["penetrance"] == fx["dm_penetrance"]).all()
    assert (h.elim_masks() == fx["elim"]).all()
    # same elimination order -> same types, cutset order, previous functions, children
    assert h.set_peel(fx["op_peelnode"])
    assert h.ops() == unpack_ops(fx)
    assert h.peel_cost() == int(fx["peel_cost"])
    h.close()


@needs_ref
@pytest.mark.parametrize("name", CASES)
def test_peel_search_and_random_descent_graph(name):
    from oracle import refapi, orcapi
    h = H.Host(*example_files(name), sex_linked=bool(X[name]))
    h.build_peel(20000 if name != "inbred" else 200000, seed=11)
    fx = golden(name)
    if name != "inbred":
        assert h.peel_cost() <= int(fx["peel_cost"])          # the small examples have a unique optimum cost
    else:
        assert h.peel_cost() <= 2 * int(fx["peel_cost"])      # random search: same ballpark as the reference's
    # the reference accepts the order and derives the same operations from it
    r = refapi.Ref(*example_files(name), sex_linked=bool(X[name]))
    assert r.set_peel([o["peelnode"] for o in h.ops()])
    assert r.ops() == h.ops()
    # random start graph from genotype elimination is legal under the reference's likelihood
    dg = h.random_descentgraph(seed=3)
    assert set(np.unique(dg)) <= {0, 1}
    r.dg_set(dg)
    assert r.dg_likelihood() > -1e300
    # and the oracle peels it with positive likelihood at every locus
    orc = orcapi.Problem(h.problem_dict())
    for l in range(h.M):
        assert orc.ls_forward(dg, l)[0] > 0.0
    r.close()
    h.close()


@needs_ref
def test_illegal_orders_and_bad_files_are_refused(tmp_path):
    h = H.Host(*example_files("loop"))
    seq = list(range(h.N))
    seq[1] = 0
    assert not h.set_peel(seq)                                           # repeats a node
    h.close()
    bad = tmp_path / "bad.ped"
    ped, mapf, dat = example_files("loop")
    lines = open(ped).read().splitlines()
    lines[3] = lines[3].replace(" 1   2 ", " 1   77 ", 1)                # mother does not exist
    bad.write_text("\n".join(lines) + "\n")
    with pytest.raises(RuntimeError):
        H.Host(str(bad), mapf, dat)


@needs_ref
def test_results_writer_format(tmp_path):
    """marker<TAB>position<TAB>lod header, one line per marker and per position (linkage_writer.cc:52-84)"""
    h = H.Host(*example_files("loop"))
    lod = np.arange((h.M - 1) * h.nlod, dtype=np.float64) / 10.0
    out = tmp_path / "swiftlink.out"
    assert h.write_results(str(out), lod)
    rows = out.read_text().splitlines()
    assert rows[0] == "marker\tposition\tlod"
    assert len(rows) == 1 + h.M + (h.M - 1) * h.nlod
    assert rows[1].split("\t") == ["rs1", "10"]
    first = rows[2].split("\t")
    assert first[0] == "-" and float(first[2]) == 0.0
    assert abs(float(rows[3].split("\t")[2]) - 0.1) < 1e-12
    assert rows[-1].split("\t")[0] == "rs3"
    h.close()


@pytest.mark.parametrize("name", CASES)
def test_flattened_problem_accepted(name):
    """the fixture problem passes host-side plan validation with the reference's prior quirk on"""
    st = capi.plan_validate(problem(name))
    assert st["n_ops"] == int(golden(name)["N"])
