"""Builds tests/emu/libslk_emu.so: the sequential CPU emulation of the peel kernels' per-thread code.
Test infrastructure only (see slk_emu.cc)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "swiftlink_b200", "csrc")
LIB = os.path.join(HERE, "libslk_emu.so")


def build(force=False):
    src = [os.path.join(HERE, "slk_emu.cc"), os.path.join(CSRC, "slk_plan.cc")]
    dep = src + [os.path.join(CSRC, f) for f in ("slk_peel.h", "slk_types.h", "slk_plan.h", "slk_philox.cuh", "slk_geometry.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(f) <= os.path.getmtime(LIB) for f in dep):
        return LIB
    # -ffp-contract=off: no fused multiply-add, as the kernels are compiled with -fmad=false
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                           "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", LIB] + src)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
