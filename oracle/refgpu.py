"""The reference's OWN GPU path for LOD scoring, as the same-box baseline of slk_lodscore_kernel -- TEST INFRASTRUCTURE ONLY.

oracle/_ref/libswiftref_gpu.so (`make -C oracle refgpu`) is the UNMODIFIED reference built the way its Makefile.cuda
builds the `-g` binary -- gpu_lodscores.cc + cuda_linkage.cu (which #includes cuda_common.cu and cuda_lodscore.cu)
-- but for sm_100a, plus the dump driver.  GPULodscores::calculate / get_results (gpu_lodscores.cc:598-637) drive
lodscore_kernel (cuda_lodscore.cu:389-467).

The reference's destructor calls cudaDeviceReset() and its runtime is the shared cudart, ours the static one, so the
measurement runs in a process of its own:

    python -m oracle.refgpu --case east --reps 20 --out /tmp/east_refgpu.npz

writes the descent graph it scored, the elimination order, the normalised LOD table lodscore_kernel produced and the
seconds per scoring pass; bench.py (`--config east|loop`) and tests/test_gpu_vs_reference_gpu.py read that file and run
slk_lodscore_kernel on the same graph with the same elimination order.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GPU_LIB = os.path.join(HERE, "_ref", "libswiftref_gpu.so")


def available():
    return os.path.exists(GPU_LIB)


def run_in_subprocess(case, reps, out, files=None, sex_linked=False, timeout=600):
    """returns the dict of arrays the child wrote, or raises RuntimeError with the child's stderr"""
    cmd = [sys.executable, "-m", "oracle.refgpu", "--case", case, "--reps", str(reps), "--out", out]
    if files:
        cmd += ["--files"] + list(files)
    if sex_linked:
        cmd += ["--sex-linked"]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0 or not os.path.exists(out):
        raise RuntimeError("reference GPU path failed (rc %d): %s" % (p.returncode, (p.stderr or p.stdout)[-600:]))
    return dict(np.load(out, allow_pickle=False))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="east")
    ap.add_argument("--files", nargs=3, default=None, help="ped map dat (default: the reference's example of that name)")
    ap.add_argument("--sex-linked", action="store_true")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--peel-iterations", type=int, default=100000)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import refapi
    refapi.LIB_PATH = GPU_LIB                      # the same driver entry points + ref_gpu_lod_bench
    L = refapi.lib()
    L.ref_gpu_lod_bench.restype = C.c_double
    refapi.set_threads(1)
    refapi.seed(20261017)
    files = args.files or refapi.example(args.case)
    r = refapi.Ref(*files, sex_linked=args.sex_linked, lodscores=5)
    r.build_peel(args.peel_iterations)
    order = np.array([op["peelnode"] for op in r.ops()], np.uint32)
    r.sequential_imputation(10)
    dg = r.dg_get()
    # the CPU peeler on the same graph (one thread), for the record
    t_cpu = r.bench_lodpasses(3)
    # ... and its table: Peeler::process per interval, one pass, normalised as LODscores::get does (lod_score.h:86-88)
    tp = r.calc_trait_prob()
    lod_cpu = np.array([(r.lod_interval(l)[1] - tp) / np.log(10.0) for l in range(r.M - 1)])
    lod = np.zeros((r.M - 1) * r.nlod)
    setup = C.c_double(0)
    threads = C.c_int(0)
    secs = float(L.ref_gpu_lod_bench(r.h, int(args.reps), lod.ctypes.data_as(C.POINTER(C.c_double)), C.byref(setup), C.byref(threads)))
    np.savez(args.out, dg=dg, order=order, lod=lod.reshape(r.M - 1, r.nlod), lod_cpu=lod_cpu, secs_per_pass=secs, setup_s=setup.value,
             block_threads=threads.value, positions=(r.M - 1) * r.nlod, cpu_secs_per_pass_1thread=t_cpu,
             trait_prob=tp, N=r.N, M=r.M)
    print(json.dumps(dict(case=args.case, secs_per_pass=secs, positions=(r.M - 1) * r.nlod,
                          trait_positions_per_s=(r.M - 1) * r.nlod / secs, setup_s=setup.value, block_threads=threads.value)))
    sys.stdout.flush()
    os._exit(0)                                     # skip ~GPULodscores / static destructors (cudaDeviceReset)


if __name__ == "__main__":
    main()
