mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_job.py tests/test_gpu_vs_reference_gpu.py tests/test_gpu_elod.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2g_tests.log 2>&1
timeout 300 python tools/profile_target.py --sweeps 2 --lod 0 --msweeps 3 --trace > gpurun_out/r2g_trace.log 2>&1
SLK_MS_TIMELINE=1 timeout 200 python tools/profile_target.py --sweeps 1 --lod 0 --msweeps 1 > gpurun_out/r2g_timeline.log 2>&1
cat gpurun_out/r2g_tests.log; grep -v "^{" gpurun_out/r2g_trace.log | cut -c1-400
