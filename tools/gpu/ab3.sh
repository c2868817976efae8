mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3)
run() { echo "== $*"; env "$@" timeout 200 python tools/profile_target.py --sweeps 30 --lod 3 --time --trace 2>&1 | tail -2; }
run SLK_LS_TEAM=64 SLK_LOD_TEAM=64 SLK_CTA_THREADS=512
run SLK_LS_TEAM=64 SLK_LOD_TEAM=64 SLK_CTA_THREADS=384
run SLK_LS_TEAM=32 SLK_LOD_TEAM=32 SLK_CTA_THREADS=384
run SLK_LS_TEAM=32 SLK_LOD_TEAM=32 SLK_CTA_THREADS=320
