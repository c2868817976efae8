python -m pytest tests/test_gpu_msampler.py tests/test_gpu_edge_cases.py -x -q 2>&1 | tail -2
for v in "SLK_NO_PDL=1" "SLK_MS_NO_PREFIX=1" "SLK_X=1"; do
echo $v; env $v python tools/profile_target.py --msweeps 5 2>&1 | grep "M-sweep"
done
python tools/profile_target.py --msweeps 3 --trace 2>&1 | grep -A14 "chain kernel CTA 0"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --in-flight 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['derived']['m_sweep_ms'], d['roofline']['launch_ms'], d['roofline']['ms_chain_kernel']['launch_ms'])"
