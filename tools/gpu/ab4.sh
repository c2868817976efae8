mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 200 python tools/profile_target.py --sweeps 40 --lod 3 --time 2>&1 | tail -1; }
for v in v000 v100 v010 v001 v000; do
  run SLK_LIB=$PWD/swiftlink_b200/libslk_$v.so SLK_LS_TEAM=64 SLK_LOD_TEAM=64 SLK_CTA_THREADS=512
  run SLK_LIB=$PWD/swiftlink_b200/libslk_$v.so SLK_LS_TEAM=32 SLK_LOD_TEAM=32 SLK_CTA_THREADS=384
done
