P="python tools/profile_target.py --sweeps 2 --lod 1 --msweeps 1"
for k in slk_ms_step:50 slk_ms_chain:50 slk_lsampler:3 slk_lodscore:0 slk_ms_likelihood:1; do
  name=${k%%:*}; skip=${k##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -f -o gpurun_out/prof_r1b_$name $P > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
done
SLK_BENCH_SCORING_PERIOD=6 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_r1b.csv
ls -la gpurun_out/*.ncu-rep
