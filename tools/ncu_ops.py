"""Dynamic opcode histogram + headline counters of one kernel in an ncu report.

    python tools/ncu_ops.py gpurun_out/x.ncu-rep <units in the launch>
"""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep, units = sys.argv[1], float(sys.argv[2])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, vals = rows[0], rows[2]
    want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size"]
    for h, v in zip(hdr, vals):
        if h in want:
            print("%-90s %s" % (h, v))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    ia, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    tot = 0
    byop, smp = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) <= ia:
            continue
        try:
            n, sm = int(r[ia]), int(r[ismp])
        except ValueError:
            continue
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[isrc].strip())
        op = m.group(2) if m else r[isrc][:8]
        byop[op] += n; smp[op] += sm; tot += n
    print("warp instructions %d, per unit %.0f" % (tot, tot / units))
    for op, n in byop.most_common(24):
        print("  %-8s %6.2f%%  %9.1f /unit   stall samples %5.1f%%" % (op, 100.0 * n / tot, n / units, 100.0 * smp[op] / max(1, sum(smp.values()))))


if __name__ == "__main__":
    main()
