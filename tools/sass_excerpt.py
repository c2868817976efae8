"""profiles/<tag>_sass_excerpt.txt: the SASS lines that show the bulk copy (UBLKCP) + mbarrier (SYNCS) staging of the
peel program, the programmatic-dependent-launch instructions of the M-sampler kernels and the FP64 / cluster
instructions, per kernel, from the library as built.

    python tools/sass_excerpt.py r2
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"\b(UBLKCP|SYNCS|ACQBULK|PREEXIT|UCGABAR|CCTL|LDGSTS|DMUL|DADD|DFMA|LDG|STG|LDS|STS|BAR|WARPSYNC|SHFL)\b")
SHOW = ("UBLKCP", "SYNCS", "ACQBULK", "PREEXIT", "UCGABAR")


def main():
    tag = sys.argv[1]
    lib = os.path.join(ROOT, "swiftlink_b200", "libswiftlink_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    out = ["SASS excerpt of swiftlink_b200/libswiftlink_b200.so (cuobjdump -sass; sm_100a), production instantiations.",
           "UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier ops, ACQBULK = griddepcontrol.wait, PREEXIT = griddepcontrol.launch_dependents,",
           "UCGABAR = cluster barrier.  Counts are static instruction counts of the kernel.", ""]
    fn, counts, shown = None, None, None
    keep = ("slk_lsampler_kernelILi32ELi384ELb0", "slk_lodscore_kernelILi96ELi576ELb0", "slk_ms_step_kernelILb0", "slk_ms_chain_kernel",
            "slk_ms_likelihood_kernelILb0")

    def flush():
        if fn and any(k in fn for k in keep):
            out.append("== %s" % fn)
            out.append("   " + ", ".join("%s %d" % kv for kv in sorted(counts.items())))
            out.extend(shown[:14])
            out.append("")
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            flush()
            fn, counts, shown = m.group(1), collections.Counter(), []
            continue
        if fn is None:
            continue
        m = PAT.search(ln)
        if m:
            counts[m.group(1)] += 1
            if m.group(1) in SHOW:
                shown.append("   " + ln.strip()[:110])
    flush()
    path = os.path.join(ROOT, "profiles", "%s_sass_excerpt.txt" % tag)
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:60]))


if __name__ == "__main__":
    main()
