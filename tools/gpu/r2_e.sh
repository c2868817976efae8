mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_msampler.py tests/test_gpu_edge_cases.py tests/test_gpu_job.py tests/test_gpu_vs_reference_gpu.py -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r2e_tests.log 2>&1
for i in 1 2; do
SLK_MS_RUN_AHEAD=1 timeout 200 python tools/profile_target.py --sweeps 2 --lod 0 --msweeps 10 2>&1 | grep "M-sweep" | sed 's/^/ahead1: /'
timeout 200 python tools/profile_target.py --sweeps 2 --lod 0 --msweeps 10 2>&1 | grep "M-sweep" | sed 's/^/ahead2: /'
done > gpurun_out/r2e_ab.log 2>&1
for c in east loop xlinked; do
  timeout 900 python bench.py --config $c --steps 3 --warmup 1 > gpurun_out/r2e_bench_$c.json 2> gpurun_out/r2e_bench_$c.err
done
cat gpurun_out/r2e_tests.log gpurun_out/r2e_ab.log
for c in east loop xlinked; do head -c 300 gpurun_out/r2e_bench_$c.json; echo; tail -2 gpurun_out/r2e_bench_$c.err; done
