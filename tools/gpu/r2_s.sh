mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2s_bench_8gpu.json 2> gpurun_out/r2s_bench_8gpu.err
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_bench_1gpu.json 2> gpurun_out/r2s_bench_1gpu.err
wc -l gpurun_out/r2s_bench_8gpu.json gpurun_out/r2s_bench_1gpu.json
head -c 250 gpurun_out/r2s_bench_8gpu.json; echo; tail -2 gpurun_out/r2s_bench_8gpu.err | cut -c1-200; head -c 250 gpurun_out/r2s_bench_1gpu.json; echo
