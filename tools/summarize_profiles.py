"""Writes profiles/<tag>_summary.md from an ncu launch list (CSV) and ncu --set full reports.

    python tools/summarize_profiles.py TAG launches.csv sampler.ncu-rep lod.ncu-rep
"""
import csv
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
       "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_elapsed",
       "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__icc_request_hit_rate.pct",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    res = {"kernel": vals[hdr.index("Kernel Name")]}
    for m in RAW:
        if m in hdr:
            i = hdr.index(m)
            res[m] = (vals[i], units[i])
    return res


def stalls(rep):
    out = subprocess.run([sys.executable, "tools/ncu_lines.py", rep, "12"], capture_output=True, text=True).stdout
    return out


SKIP = int(__import__("os").environ.get("SLK_SUMMARY_SKIP", "0"))       # leading launches that belong to the set-up


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [r for r in rows if "Kernel Name" in r][0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = {}
    seen = 0
    for r in rows:
        if len(r) == len(hdr) and r is not hdr and r[v].replace(".", "").isdigit():
            seen += 1
            if seen <= SKIP:
                continue
            name = r[k].split("(")[0]
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += float(r[v])
    return agg


def main():
    tag, lcsv, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    lines = ["# ncu summary %s" % tag, "",
             "Launch list: `SLK_BENCH_SCORING_PERIOD=6 ncu --metrics gpu__time_duration.sum --clock-control none ... "
             "python bench.py --steps 2 --warmup 1 --no-cpu-baseline` (a step of 6 iterations + 1 scoring pass instead of 100 + 1 "
             "so that the whole run fits under ncu; the first %d launches -- sequential imputation and burn-in -- are left out; "
             "per-launch times are cold-cache and serialised: compare shares, not absolutes)." % SKIP, "",
             "| kernel | launches | total ms | share | avg ms |", "|---|---:|---:|---:|---:|"]
    agg = launches(lcsv)
    tot = sum(a[1] for a in agg.values())
    for name, (n, ns) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append("| `%s` | %d | %.2f | %.1f %% | %.3f |" % (name, n, ns / 1e6, 100 * ns / tot, ns / n / 1e6))
    for rep in reps:
        m = raw_metrics(rep)
        lines += ["", "## `%s` (%s, `ncu --set full --clock-control none --import-source on`)" % (m.pop("kernel"), rep), "",
                  "| metric | value |", "|---|---|"]
        for k_, (val, unit) in m.items():
            lines.append("| %s | %s %s |" % (k_, val, unit))
        lines += ["", "```", stalls(rep).rstrip(), "```"]
    open("profiles/%s_summary.md" % tag, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:14]))


if __name__ == "__main__":
    main()
