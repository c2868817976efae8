set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:slk_lsampler -s 6 -c 1 -o gpurun_out/ls_v2a python tools/profile_target.py --markers 4000 --sweeps 4 --lod 0 > gpurun_out/ncu_ls.log 2>&1
tail -5 gpurun_out/ncu_ls.log
ls -la gpurun_out/
