"""Per-source-line summary of an ncu report (needs -lineinfo and --import-source on).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    data = []
    for r in rows:
        if len(r) > 10 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[2] != "-":
            continue                       # keep the per-source-line aggregate rows only
        def col(name):
            v = r[hdr.index(name)]
            try:
                return int(v)
            except ValueError:
                return 0
        data.append(dict(line=r[0], src=r[1].strip(), samples=col("# Samples"), inst=col("Instructions Executed"),
                         wf=col("L1 Wavefronts Shared"), wfi=col("L1 Wavefronts Shared Ideal"),
                         long_sb=col("stall_long_sb"), short_sb=col("stall_short_sb"), barrier=col("stall_barrier"),
                         mio=col("stall_mio"), wait=col("stall_wait"), math=col("stall_math"), noinst=col("stall_no_inst"),
                         branch=col("stall_branch_resolving")))
    ts = float(sum(d["samples"] for d in data)) or 1.0
    ti = float(sum(d["inst"] for d in data)) or 1.0
    print("total samples %d, warp instructions %d, shared wavefronts %d (ideal %d)" %
          (ts, ti, sum(d["wf"] for d in data), sum(d["wfi"] for d in data)))
    for key in ("long_sb", "short_sb", "barrier", "mio", "wait", "math", "noinst", "branch"):
        print("  stall %-9s %5.1f%%" % (key, 100.0 * sum(d[key] for d in data) / ts))
    print("--- top %d lines by stall samples" % top)
    for d in sorted(data, key=lambda d: -d["samples"])[:top]:
        print("%5s %5.1f%% smp %5.1f%% inst  wf %10d/%10d  %s" % (d["line"], 100 * d["samples"] / ts, 100 * d["inst"] / ti,
                                                                 d["wf"], d["wfi"], d["src"][:100]))


if __name__ == "__main__":
    main()
