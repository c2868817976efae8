mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_elod.py tests/test_gpu_msampler.py -m gpu -x -q -k "lod or trait or elod or bench or full_size or edge" 2>&1 | tail -5) > gpurun_out/r2n_tests.log 2>&1
timeout 300 python tools/profile_target.py --sweeps 4 --lod 4 --time 2>&1 | grep "sweep ms" >> gpurun_out/r2n_tests.log
cat gpurun_out/r2n_tests.log
