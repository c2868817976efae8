python -m pytest tests/test_gpu_msampler.py tests/test_gpu_edge_cases.py -x -q 2>&1 | tail -2
SLK_MS_TIMELINE=1 python tools/profile_target.py --msweeps 1 2>&1 | grep -E "pair|M-sweep" | tail -9
for v in "SLK_MS_CHAIN_EXCLUSIVE=0" "SLK_X=1"; do
echo $v; env $v python tools/profile_target.py --msweeps 5 2>&1 | grep "M-sweep"
done
