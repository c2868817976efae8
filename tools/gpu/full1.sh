mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err
tail -c 3000 gpurun_out/bench_v2.json; tail -3 gpurun_out/bench_v2.err
