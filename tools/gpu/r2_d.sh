mkdir -p gpurun_out
for ra in 1 2; do
echo "== run-ahead $ra"
SLK_MS_RUN_AHEAD=$ra timeout 600 python -m pytest tests/test_gpu_msampler.py -m gpu -q -k "sweeps_match_oracle or bench_pedigree_sweeps or replicates" 2>&1 | tail -8
done > gpurun_out/r2d_tests.log 2>&1
echo "== run-ahead 2, exclusive chain kernel" >> gpurun_out/r2d_tests.log
SLK_MS_CHAIN_EXCLUSIVE=1 timeout 600 python -m pytest tests/test_gpu_msampler.py -m gpu -q -k "sweeps_match_oracle" 2>&1 | tail -8 >> gpurun_out/r2d_tests.log
echo "== run-ahead 2, no rec" >> gpurun_out/r2d_tests.log
SLK_MS_NO_REC=1 timeout 600 python -m pytest tests/test_gpu_msampler.py -m gpu -q -k "sweeps_match_oracle" 2>&1 | tail -8 >> gpurun_out/r2d_tests.log
cat gpurun_out/r2d_tests.log
