"""bench.py / run.py command lines that need no GPU: the reference arm of the configuration that has none prints the
contract's `unavailable` line; the multi-GPU entry point parses the `swift` options it mirrors."""
import json
import os
import subprocess
import sys

from common import ROOT


def test_reference_arm_of_c4_is_declared_unavailable():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c4"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-1000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and "unavailable" in d


def test_run_entry_point_options():
    from swiftlink_b200 import run
    a = run.parse(["-p", "x.ped", "-m", "x.map", "-d", "x.dat", "-R", "8", "-M", "-z", "3", "-y", "5", "-b", "10", "-i", "20",
                   "-t", "1.0,0.9,0.8", "-X"])
    assert (a.runs, a.mcmcmc, a.chains, a.exchangeperiod, a.burnin, a.iterations, a.sexlinked) == (8, True, 3, 5, 10, 20, True)
    assert [float(x) for x in a.temperatures.split(",")] == [1.0, 0.9, 0.8]
    from swiftlink_b200 import dist as sdist
    assert sdist.chain_placement(8, 3) == [[0, 3, 6], [1, 4, 7], [2, 5]]


def test_small_config_descriptions_name_the_reference_commands():
    sys.path.insert(0, ROOT)
    import bench
    for name, cfg in bench.SMALL_CONFIGS.items():
        d = bench.small_config_desc(name)
        assert cfg["cli"] in d["workload"] and d["replicates"] == cfg["runs"]
    assert bench.SMALL_CONFIGS["loop"]["runs"] == 10 and bench.SMALL_CONFIGS["xlinked"]["sex_linked"]
