mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:slk_ms_step -s 20 -c 1 -o gpurun_out/ms_step_r2 python tools/profile_target.py --sweeps 1 --lod 0 --msweeps 1 > gpurun_out/ncu_ms.log 2>&1
tail -3 gpurun_out/ncu_ms.log
