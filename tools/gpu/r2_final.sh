# final evidence of the round: tests, bench lines of every configuration, ncu launch list and full captures
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/final_tests.log 2>&1
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
for c in east loop xlinked; do
  timeout 900 python bench.py --config $c --steps 3 --warmup 1 > gpurun_out/final_bench_$c.json 2> gpurun_out/final_bench_$c.err
done
timeout 900 python bench.py --config c4 --steps 3 --warmup 1 > gpurun_out/final_bench_c4_1gpu.json 2> gpurun_out/final_bench_c4_1gpu.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
SLK_BENCH_SCORING_PERIOD=6 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 > gpurun_out/final_ncu_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:slk_ms_step -s 20 -c 1 -o gpurun_out/final_ms_step python tools/profile_target.py --sweeps 1 --lod 0 --msweeps 1 > gpurun_out/final_ncu_ms_step.log 2>&1
timeout 600 $NCU -k regex:slk_ms_chain -s 20 -c 1 -o gpurun_out/final_ms_chain python tools/profile_target.py --sweeps 1 --lod 0 --msweeps 1 > gpurun_out/final_ncu_ms_chain.log 2>&1
timeout 600 $NCU -k regex:slk_ms_likelihood -s 1 -c 1 -o gpurun_out/final_ms_likelihood python tools/profile_target.py --sweeps 1 --lod 0 --msweeps 1 > gpurun_out/final_ncu_ms_lik.log 2>&1
timeout 600 $NCU -k regex:slk_lsampler -s 6 -c 1 -o gpurun_out/final_ls python tools/profile_target.py --sweeps 4 --lod 0 > gpurun_out/final_ncu_ls.log 2>&1
timeout 900 $NCU -k regex:slk_lodscore -s 0 -c 1 -o gpurun_out/final_lod python tools/profile_target.py --sweeps 1 --lod 1 > gpurun_out/final_ncu_lod.log 2>&1
cat gpurun_out/final_tests.log gpurun_out/final_smoke.log
for f in final_bench final_bench_reference final_bench_east final_bench_loop final_bench_xlinked final_bench_c4_1gpu; do echo "== $f"; head -c 260 gpurun_out/$f.json; echo; tail -2 gpurun_out/$f.err; done
ls -la gpurun_out/final_*
