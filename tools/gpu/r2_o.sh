mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_msampler.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -4) > gpurun_out/r2o_tests.log 2>&1
for i in 1 2 3; do timeout 300 python tools/profile_target.py --sweeps 2 --lod 0 --msweeps 10 2>&1 | grep "M-sweep"; done >> gpurun_out/r2o_tests.log 2>&1
timeout 300 python tools/profile_target.py --sweeps 1 --lod 0 --msweeps 2 --trace 2>&1 | grep "chain kernel" | cut -c1-300 >> gpurun_out/r2o_tests.log
timeout 300 python tools/profile_target.py --sweeps 1 --lod 0 --timeline 2>&1 | grep "pair\|CTAs" | cut -c1-300 >> gpurun_out/r2o_tests.log
cat gpurun_out/r2o_tests.log
