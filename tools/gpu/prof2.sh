mkdir -p gpurun_out
timeout 200 python tools/profile_target.py --sweeps 30 --lod 3 --time --trace 2>&1 | tail -2
SLK_LS_TEAM=32 SLK_LOD_TEAM=32 timeout 200 python tools/profile_target.py --sweeps 30 --lod 3 --time 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:slk_lsampler -s 6 -c 1 -o gpurun_out/ls_v2b python tools/profile_target.py --markers 4000 --sweeps 4 --lod 0 > gpurun_out/ncu_ls.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:slk_lodscore -s 0 -c 1 -o gpurun_out/lod_v2b python tools/profile_target.py --markers 2000 --sweeps 1 --lod 1 > gpurun_out/ncu_lod.log 2>&1
ls -la gpurun_out
