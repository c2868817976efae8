mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2r_bench_8gpu.json 2> gpurun_out/r2r_bench_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --config c4 --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2r_c4_8gpu.json 2> gpurun_out/r2r_c4_8gpu.err
head -c 300 gpurun_out/r2r_bench_8gpu.json; echo; tail -2 gpurun_out/r2r_bench_8gpu.err | cut -c1-200; head -c 300 gpurun_out/r2r_c4_8gpu.json; echo; tail -2 gpurun_out/r2r_c4_8gpu.err | cut -c1-200
