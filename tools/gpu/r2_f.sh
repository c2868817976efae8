mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r2f_tests.log 2>&1
for i in 1 2; do
SLK_MS_RUN_AHEAD=1 timeout 200 python tools/profile_target.py --sweeps 2 --lod 0 --msweeps 10 2>&1 | grep "M-sweep" | sed 's/^/ahead1: /'
timeout 200 python tools/profile_target.py --sweeps 2 --lod 0 --msweeps 10 2>&1 | grep "M-sweep" | sed 's/^/ahead2: /'
done > gpurun_out/r2f_ab.log 2>&1
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_msampler.py -m gpu -x -q -k "sweeps or replicates or full_size" 2>&1 | tail -2; done > gpurun_out/r2f_repeat.log 2>&1
cat gpurun_out/r2f_tests.log gpurun_out/r2f_ab.log gpurun_out/r2f_repeat.log
