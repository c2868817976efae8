"""The `swift` command line (swiftlink_b200/swift): flags, refusals and the output file."""
import os
import subprocess

import numpy as np
import pytest

from common import ROOT, golden, ref_available

SWIFT = os.path.join(ROOT, "swiftlink_b200", "swift")
needs_ref = pytest.mark.skipif(not ref_available(), reason="example inputs live in oracle/_ref/examples")


def test_cli_builds_and_prints_usage():
    from swiftlink_b200 import build
    build.build()
    out = subprocess.run([SWIFT, "-h"], capture_output=True, text=True)
    assert out.returncode == 0
    for flag in ("--pedigree", "--map", "--dat", "--iterations", "--burnin", "--scoringperiod", "--lodscores",
                 "--runs", "--sexlinked", "--affectedonly", "--peelseqiter"):
        assert flag in out.stderr


@needs_ref
def test_cli_refuses_what_it_cannot_do():
    from oracle import refapi
    ped, mapf, dat = refapi.example("loop")
    out = subprocess.run([SWIFT, "-p", ped, "-m", mapf], capture_output=True, text=True)
    assert out.returncode != 0 and "required" in out.stderr
    out = subprocess.run([SWIFT, "-p", ped, "-e", "-k", "0.1,0.2"], capture_output=True, text=True)
    assert out.returncode != 0 and "penetrance requires 3 floats" in out.stderr
    out = subprocess.run([SWIFT, "-p", ped, "-m", mapf, "-d", dat, "-l", "1.5"], capture_output=True, text=True)
    assert out.returncode != 0 and "between 0.0 and 1.0" in out.stderr


@pytest.mark.gpu
@needs_ref
def test_cli_end_to_end_east(tmp_path):
    """swift -p east.ped -m east.map -d east.dat (default sampler mix, short run, with the CODA trace) writes the
    reference's marker / position / lod table, the reference's P(T) and the trace files"""
    from oracle import refapi
    ped, mapf, dat = refapi.example("east")
    out_file = tmp_path / "swiftlink.out"
    out = subprocess.run([SWIFT, "-p", ped, "-m", mapf, "-d", dat, "-b", "300", "-i", "600", "-s", "5",
                          "-q", "20000", "-R", "2", "-T", "-P", str(tmp_path / "tr"), "-o", str(out_file)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "P(T) = -9.31939" in out.stdout                          # BASELINE.md known-answer scalar
    rows = out_file.read_text().splitlines()
    fx = golden("east")
    M, n = int(fx["M"]), int(fx["nlod"])
    assert rows[0] == "marker\tposition\tlod"
    assert len(rows) == 1 + M + (M - 1) * n
    assert rows[1].split("\t")[0] == "marker1" and abs(float(rows[1].split("\t")[1]) - 10.0) < 1e-9
    lods = np.array([float(r.split("\t")[2]) for r in rows[1:] if r.startswith("-")])
    assert lods.shape == ((M - 1) * n,) and np.isfinite(lods).all()
    assert -5.0 < lods.mean() < 8.0
    for run in (0, 1):
        tr = [p for p in tmp_path.iterdir() if p.name.startswith("tr.ped") and p.name.endswith(".run%d" % run)]
        assert len(tr) == 1
        lines = tr[0].read_text().splitlines()
        assert lines[0] == "iteration likelihood" and len(lines) == 1 + 60          # iterations 300..899 every 10th
        assert all(float(ln.split("\t")[1]) < 0.0 for ln in lines[1:])


@pytest.mark.gpu
@needs_ref
def test_cli_elod():
    """swift -p loop.ped --elod -u 100000: the reference's report format and an estimate in its Monte Carlo band"""
    from oracle import refapi
    ref = np.load(os.path.join(ROOT, "tests", "golden", "elod_ref.npz"))
    out = subprocess.run([SWIFT, "-p", refapi.example("loop")[0], "-e", "-u", "100000", "-q", "20000"], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "ELOD parameters:" in out.stderr and "replicates = 100000" in out.stderr
    line = [ln for ln in out.stderr.splitlines() if ln.startswith("ELOD = ")]
    assert len(line) == 1
    assert abs(float(line[0].split("=")[1]) - ref["loop"].mean()) < 0.01
