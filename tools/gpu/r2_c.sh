mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_msampler.py tests/test_gpu_edge_cases.py tests/test_gpu_vs_reference_gpu.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2c_tests.log 2>&1
for i in 1 2; do
SLK_MS_RUN_AHEAD=1 timeout 200 python tools/profile_target.py --sweeps 2 --lod 3 --msweeps 10 --time 2>&1 | grep "M-sweep\|sweep ms" | sed 's/^/ahead1: /'
timeout 200 python tools/profile_target.py --sweeps 2 --lod 3 --msweeps 10 --time 2>&1 | grep "M-sweep\|sweep ms" | sed 's/^/ahead2: /'
SLK_MS_CHAIN_EXCLUSIVE=1 timeout 200 python tools/profile_target.py --sweeps 2 --lod 0 --msweeps 10 2>&1 | grep "M-sweep" | sed 's/^/ahead2 exclusive: /'
done > gpurun_out/r2c_ab.log 2>&1
SLK_MS_TIMELINE=1 timeout 200 python tools/profile_target.py --sweeps 1 --lod 0 --msweeps 1 > gpurun_out/r2c_timeline.log 2>&1
for c in east loop xlinked; do
  timeout 900 python bench.py --config $c --steps 3 --warmup 1 > gpurun_out/r2c_bench_$c.json 2> gpurun_out/r2c_bench_$c.err
done
cat gpurun_out/r2c_tests.log gpurun_out/r2c_ab.log
grep "pair\|step kernel CTAs" gpurun_out/r2c_timeline.log | head -12
for c in east loop xlinked; do head -c 200 gpurun_out/r2c_bench_$c.json; echo; tail -2 gpurun_out/r2c_bench_$c.err; done
