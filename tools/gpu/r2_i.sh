mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --config c4 --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2i_c4_2gpu.json 2> gpurun_out/r2i_c4_2gpu.err
head -c 400 gpurun_out/r2i_c4_2gpu.json; tail -5 gpurun_out/r2i_c4_2gpu.err
