mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2q_bench_2gpu.json 2> gpurun_out/r2q_bench_2gpu.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --config c4 --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2q_c4_2gpu.json 2> gpurun_out/r2q_c4_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --ref-step-seconds 2 > gpurun_out/r2q_ref_2gpu.json 2> gpurun_out/r2q_ref_2gpu.err
head -c 300 gpurun_out/r2q_bench_2gpu.json; echo; tail -3 gpurun_out/r2q_bench_2gpu.err; head -c 300 gpurun_out/r2q_c4_2gpu.json; echo; tail -2 gpurun_out/r2q_c4_2gpu.err | cut -c1-200; head -c 200 gpurun_out/r2q_ref_2gpu.json; tail -2 gpurun_out/r2q_ref_2gpu.err
