mkdir -p gpurun_out
for cfg in "64 384" "64 320" "64 256" "96 576" "96 480" "96 384" "128 512" "192 576" "256 768"; do
  set -- $cfg
  SLK_LOD_TEAM=$1 SLK_LOD_CTA_THREADS=$2 timeout 300 python tools/profile_target.py --sweeps 2 --lod 4 --time 2>&1 | grep "sweep ms\|lod_smem" | sed -e "s/.*'lod_smem_doubles': \([0-9.]*\).*'lod_cta_smem': \([0-9.]*\).*/smem_doubles \1 cta_smem \2/" | sed "s/^/[lod team $1 cta $2] /"
done > gpurun_out/r2m_ab.log 2>&1
cat gpurun_out/r2m_ab.log
