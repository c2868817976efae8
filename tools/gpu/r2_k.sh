mkdir -p gpurun_out
for v in "" bwd4 wave1 wave1b; do
  for i in 1 2; do
    if [ -z "$v" ]; then lib=""; else lib="$PWD/swiftlink_b200/libslk_$v.so"; fi
    SLK_LIB=$lib timeout 300 python tools/profile_target.py --sweeps 10 --lod 3 --time 2>&1 | grep "sweep ms" | sed "s/^/[$v] /"
  done
done > gpurun_out/r2k_ab.log 2>&1
cat gpurun_out/r2k_ab.log
