mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r2_tests.log 2>&1
for i in 1 2; do
SLK_LIB=$PWD/swiftlink_b200/libslk_oldwalk.so timeout 200 python tools/profile_target.py --sweeps 2 --lod 0 --msweeps 10 2>&1 | grep "M-sweep" | sed 's/^/old: /'
timeout 200 python tools/profile_target.py --sweeps 2 --lod 0 --msweeps 10 2>&1 | grep "M-sweep" | sed 's/^/new: /'
done > gpurun_out/r2_ab_walk.log 2>&1
timeout 600 python bench.py > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
cat gpurun_out/r2_tests.log gpurun_out/r2_ab_walk.log
head -c 600 gpurun_out/r2_bench_a.json
