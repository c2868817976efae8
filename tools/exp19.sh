python -m pytest tests/test_gpu_msampler.py tests/test_gpu_edge_cases.py -x -q 2>&1 | tail -2
python tools/profile_target.py --msweeps 5 2>&1 | grep "M-sweep"
SLK_MS_TIMELINE=1 python tools/profile_target.py --msweeps 1 2>&1 | grep -E "step kernel CTAs" | tail -2
