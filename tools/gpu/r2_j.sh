mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_elod.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2j_tests.log 2>&1
for i in 1 2; do timeout 300 python tools/profile_target.py --sweeps 10 --lod 3 --msweeps 5 --time 2>&1 | grep "M-sweep\|sweep ms"; done > gpurun_out/r2j_time.log 2>&1
timeout 300 python tools/profile_target.py --sweeps 2 --lod 0 --trace 2>&1 | grep "^stage" | cut -c1-600 >> gpurun_out/r2j_time.log
cat gpurun_out/r2j_tests.log gpurun_out/r2j_time.log
