mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_vs_reference_gpu.py tests/test_gpu_msampler.py -m gpu -x -q -k "reference_gpu or bench_pedigree_lod or full_size" 2>&1 | tail -15) > gpurun_out/r2b_tests.log 2>&1
for c in east loop xlinked; do
  timeout 900 python bench.py --config $c --steps 3 --warmup 1 > gpurun_out/r2b_bench_$c.json 2> gpurun_out/r2b_bench_$c.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slk_lsampler -s 6 -c 1 -o gpurun_out/ls_r2c python tools/profile_target.py --markers 4000 --sweeps 4 --lod 0 > gpurun_out/ncu_ls_r2c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slk_lodscore -s 0 -c 1 -o gpurun_out/lod_r2c python tools/profile_target.py --markers 2000 --sweeps 1 --lod 1 > gpurun_out/ncu_lod_r2c.log 2>&1
cat gpurun_out/r2b_tests.log
for c in east loop xlinked; do head -c 300 gpurun_out/r2b_bench_$c.json; echo; tail -3 gpurun_out/r2b_bench_$c.err; done
