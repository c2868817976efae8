mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_msampler.py tests/test_gpu_edge_cases.py tests/test_gpu_job.py -m gpu -x -q 2>&1 | tail -12) > gpurun_out/r2p_tests.log 2>&1
for i in 1 2; do
for k in 0 2 3 4 6 8; do SLK_MS_SNAPSHOT=$k timeout 300 python tools/profile_target.py --sweeps 2 --lod 0 --msweeps 10 2>&1 | grep "M-sweep" | sed "s/^/[snapshot $k] /"; done
done >> gpurun_out/r2p_tests.log 2>&1
timeout 300 python tools/profile_target.py --sweeps 1 --lod 0 --timeline 2>&1 | grep "pair\|CTAs" | cut -c1-300 >> gpurun_out/r2p_tests.log
cat gpurun_out/r2p_tests.log
