"""Top source lines of an ncu report by executed warp instructions.   python tools/ncu_topinst.py rep units [n]"""
import csv, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; data = []
for r in rows:
    if len(r) > 10 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-": continue
    try: inst = int(r[hdr.index("Instructions Executed")]); smp = int(r[hdr.index("# Samples")])
    except ValueError: continue
    data.append((r[0], r[1].strip(), inst, smp))
ti = sum(d[2] for d in data); ts = sum(d[3] for d in data)
for ln, src, inst, smp in sorted(data, key=lambda d: -d[2])[:top]:
    print("%5s %5.1f%% inst %7.0f/unit %5.1f%% smp  %s" % (ln, 100.0 * inst / ti, inst / units, 100.0 * smp / max(ts, 1), src[:105]))
