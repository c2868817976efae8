"""Writes profiles/ncu_traffic_<tag>.json -- the per-launch ncu figures bench.py quotes next to its live timings --
and profiles/<tag>_<kernel>_raw.csv (the full raw page of each capture, so the summaries can be re-derived) from
`ncu --set full` reports.

    python tools/make_traffic_json.py r2 gpurun_out/final_ms_step.ncu-rep gpurun_out/final_ls.ncu-rep ...
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    out = {"_source": "per launch, from the ncu --set full --clock-control none captures of the 200-member x 10k-SNP workload "
                      "summarised in profiles/%s_final_summary.md (raw pages: profiles/%s_*_raw.csv)" % (tag, tag),
           "_issue_active_pct": {}, "_fp64_pipe_pct": {}, "_warp_inst_per_launch": {}, "_launch_us_under_ncu": {},
           "_issue_source": "smsp__issue_active.avg.pct_of_peak_sustained_active, sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active, "
                            "smsp__inst_executed.sum of the same captures"}
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units, vals = rows[0], rows[1], rows[2]
        name = vals[hdr.index("Kernel Name")].split("<")[0].split("(")[0].replace("void ", "").strip()
        with open(os.path.join(ROOT, "profiles", "%s_%s_raw.csv" % (tag, name)), "w") as f:
            f.write(raw)

        def val(m):
            v, u = float(vals[hdr.index(m)].replace(",", "")), units[hdr.index(m)]
            scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(u, 1.0)
            return v * scale
        out[name] = int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
        out["_issue_active_pct"][name] = round(val("smsp__issue_active.avg.pct_of_peak_sustained_active"), 2)
        out["_fp64_pipe_pct"][name] = round(val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), 2)
        out["_warp_inst_per_launch"][name] = int(val("smsp__inst_executed.sum"))
        out["_launch_us_under_ncu"][name] = round(val("gpu__time_duration.sum"), 2)
    with open(os.path.join(ROOT, "profiles", "ncu_traffic_%s.json" % tag), "w") as f:
        json.dump(out, f, indent=2)
    print(json.dumps(out, indent=2))


if __name__ == "__main__":
    main()
