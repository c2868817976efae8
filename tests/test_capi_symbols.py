"""The C-ABI library loads without a GPU and exports every symbol include/swiftlink_b200.h declares."""
import ctypes
import os
import re

import pytest

from common import ROOT, CASES, problem
from swiftlink_b200 import capi, build


def header_functions():
    text = open(os.path.join(ROOT, "include", "swiftlink_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(slk_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    build.build()
    L = ctypes.CDLL(capi.LIB_PATH)
    declared = header_functions()
    assert declared, "no functions parsed from the header"
    for name in declared:
        assert hasattr(L, name), "missing export: " + name
    assert sorted(capi.SYMBOLS) == declared
    assert L.slk_abi_version() == 4


def test_no_cpu_fallback_without_device():
    """Without a usable device plan creation must fail loudly (SLK_ERR_NO_DEVICE), never compute."""
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.SlkError) as e:
        capi.Plan(problem("loop"))
    assert e.value.code == capi.ERR_NO_DEVICE


@pytest.mark.parametrize("name", CASES)
def test_plan_flattening_stats(name):
    d = problem(name)
    st = capi.plan_validate(d)
    cs = [len(o["cutset"]) for o in d["ops"]]
    assert st["n_ops"] == d["N"]
    assert st["sum_cells"] == sum(4 ** c for c in cs)
    assert st["sum_presum"] == 4 * st["sum_cells"]
    assert st["max_cutset"] == max(cs)
    # the sampler keeps every matrix (padded layout: two doubles after every 16, one-cell matrices take two)
    padded = sum(((4 ** c + 2 * (4 ** c // 16)) + 1) // 2 * 2 for c in cs)
    assert st["sum_cells"] <= st["ls_arena_doubles"] <= padded + 16
    assert st["lod_arena_doubles"] <= padded                   # trait arena reuses dead matrices
    # F_L = sum 4^(c+1) (1 + p + t) + 384 tables + 8N + 6(N-F)
    fl = 0
    tables = 0
    for o in d["ops"]:
        t = 1 if o["type"] == 1 else (len(o["children"]) if o["type"] == 2 else 0)
        tables += t
        fl += 4 ** (len(o["cutset"]) + 1) * (1 + len(o["previous"]) + t)
    fl += 64 * 6 * tables + 8 * d["N"] + 6 * (d["N"] - d["F"])
    assert st["flops_ls"] == fl


def test_plan_validation_rejects_malformed():
    d = problem("loop")
    bad = dict(d)
    bad["ops"] = [dict(o) for o in d["ops"]]
    bad["ops"][3] = dict(bad["ops"][3], previous=[7])          # consumes a later function
    with pytest.raises(capi.SlkError) as e:
        capi.plan_validate(bad)
    assert e.value.code == capi.ERR_INVALID
    bad = dict(d)
    bad["theta"] = d["theta"].copy()
    bad["theta"][0] = 0.0                                      # genetic_map.cc:48-53 refuses theta = 0
    with pytest.raises(capi.SlkError):
        capi.plan_validate(bad)
