python -m pytest tests/test_gpu_msampler.py tests/test_gpu_edge_cases.py -x -q 2>&1 | tail -2
echo "all phase 1, no PDL"; SLK_NO_PDL=1 python tools/ms_kernel_time.py 2>&1 | grep -E "kernel|cycles" | head -3
echo "all phase 0, no PDL"; SLK_NO_PDL=1 SLK_MS_DEBUG_PREV0=1 python tools/ms_kernel_time.py 2>&1 | grep -E "kernel|cycles" | head -3
SLK_MS_TIMELINE=1 python tools/profile_target.py --msweeps 1 2>&1 | grep -E "pair|M-sweep" | tail -4
python tools/profile_target.py --msweeps 5 2>&1 | grep "M-sweep"
