set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_elod.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/run1_tests.log 2>&1
cat gpurun_out/run1_tests.log
timeout 300 python tools/profile_target.py --sweeps 30 --lod 3 --time --trace > gpurun_out/run1_time.log 2>&1
cat gpurun_out/run1_time.log | tail -5
