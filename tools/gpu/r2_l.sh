mkdir -p gpurun_out
for i in 1 2; do
SLK_LOD_CTA_THREADS=384 SLK_LS_CTA_THREADS=384 timeout 300 python tools/profile_target.py --sweeps 10 --lod 3 --time 2>&1 | grep "sweep ms\|lod_cta" | cut -c1-700 | sed "s/^/[384] /"
timeout 300 python tools/profile_target.py --sweeps 10 --lod 3 --time 2>&1 | grep "sweep ms\|lod_cta"| cut -c1-700 | sed "s/^/[default] /"
done > gpurun_out/r2l_ab.log 2>&1
cat gpurun_out/r2l_ab.log
