mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -5)
run() { echo "== $*"; env "$@" timeout 200 python tools/profile_target.py --sweeps 30 --lod 3 --time --trace 2>&1 | tail -2; }
run SLK_LS_TEAM=128 SLK_LOD_TEAM=128 SLK_CTA_THREADS=384
run SLK_LS_TEAM=128 SLK_LOD_TEAM=128 SLK_CTA_THREADS=512
run SLK_LS_TEAM=128 SLK_LOD_TEAM=128 SLK_CTA_THREADS=640
run SLK_LS_TEAM=96 SLK_LOD_TEAM=96 SLK_CTA_THREADS=576
run SLK_LS_TEAM=96 SLK_LOD_TEAM=96 SLK_CTA_THREADS=384
run SLK_LS_TEAM=64 SLK_LOD_TEAM=64 SLK_CTA_THREADS=384
run SLK_LS_TEAM=64 SLK_LOD_TEAM=64 SLK_CTA_THREADS=256
run SLK_LS_TEAM=192 SLK_LOD_TEAM=192 SLK_CTA_THREADS=576
run SLK_LS_TEAM=256 SLK_LOD_TEAM=256 SLK_CTA_THREADS=768
