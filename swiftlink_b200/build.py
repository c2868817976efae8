"""Builds swiftlink_b200/libswiftlink_b200.so (CUDA kernels + C ABI + C++ host side) in-tree with
nvcc for sm_100a.  nvcc cross-compiles without a GPU, so this runs on the CPU build box too.

    python -m swiftlink_b200.build [--force]
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libswiftlink_b200.so")
SWIFT = os.path.join(HERE, "swift")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # the reference's CPU arithmetic has no fused multiply-add (g++ -O2, x86-64 baseline);
    # keeping products and sums separate makes peel matrices bit-identical to it
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    src = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cc")) +
                 [f for f in glob.glob(os.path.join(CSRC, "host", "*.cc")) if not f.endswith("swift_main.cc")])
    hdr = sorted(glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                 glob.glob(os.path.join(CSRC, "host", "*.h")) + glob.glob(os.path.join(ROOT, "include", "*.h")))
    return src, hdr


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(SWIFT):
        return True
    src, hdr = sources()
    t = min(os.path.getmtime(LIB), os.path.getmtime(SWIFT))
    return any(os.path.getmtime(f) > t for f in src + hdr + [os.path.join(CSRC, "host", "swift_main.cc")])


def build_variant(name, defs):
    """an A/B variant of the library (kernel tuning only): swiftlink_b200/libslk_<name>.so built with extra -D flags"""
    src, _ = sources()
    out = os.path.join(HERE, "libslk_%s.so" % name)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    subprocess.check_call([nvcc] + NVCC_FLAGS + list(defs) + ["-ccbin", "/usr/bin/g++", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
                           "-I", os.path.join(CSRC, "host"), "-o", out] + src)
    return out


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    src, _ = sources()
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("SLK_NVCC_DEFS", "").split()        # tuning experiments only
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-ccbin", "/usr/bin/g++", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
                                 "-I", os.path.join(CSRC, "host"), "-o", LIB] + src
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    # the `swift` command line (swiftlink_b200/swift), linked against the library next to it
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
                           "-I", os.path.join(CSRC, "host"), "-o", SWIFT, os.path.join(CSRC, "host", "swift_main.cc"),
                           "-L", HERE, "-lswiftlink_b200", "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
