# quick verification of a checkout on a B200: GPU tests, smoke, the default bench line and the reference arm
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4) > gpurun_out/verify_tests.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/verify_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/verify_bench_reference.json 2> gpurun_out/verify_bench_reference.err
cat gpurun_out/verify_tests.log gpurun_out/verify_smoke.log; wc -l gpurun_out/verify_bench.json; head -c 400 gpurun_out/verify_bench.json; echo; head -c 300 gpurun_out/verify_bench_reference.json; echo
