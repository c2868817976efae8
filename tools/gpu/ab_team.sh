mkdir -p gpurun_out
for cfg in "SLK_LS_TEAM=256 SLK_LOD_TEAM=256" "SLK_LS_TEAM=192 SLK_LOD_TEAM=192" "SLK_LS_TEAM=128 SLK_LOD_TEAM=128" "SLK_LS_TEAM=128 SLK_LOD_TEAM=128 SLK_LS_CTA_THREADS=256 SLK_LOD_CTA_THREADS=256"; do
  echo "== $cfg"; env $cfg timeout 200 python tools/profile_target.py --sweeps 30 --lod 3 --time --trace 2>&1 | tail -2
done
