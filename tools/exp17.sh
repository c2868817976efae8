python -m pytest tests/test_gpu_msampler.py tests/test_gpu_edge_cases.py -x -q 2>&1 | tail -2
python tools/profile_target.py --msweeps 5 2>&1 | grep "M-sweep"
python tools/ms_kernel_time.py 2>&1 | grep -E "kernel|cycles" | head -3
