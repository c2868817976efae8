mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_vs_reference_gpu.py tests/test_gpu_job.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -12) > gpurun_out/r2h_tests.log 2>&1
timeout 1200 python bench.py --config c4 --steps 3 --warmup 1 > gpurun_out/r2h_c4_1gpu.json 2> gpurun_out/r2h_c4_1gpu.err
cat gpurun_out/r2h_tests.log; head -c 1500 gpurun_out/r2h_c4_1gpu.json; tail -5 gpurun_out/r2h_c4_1gpu.err
