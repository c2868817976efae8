python -m pytest tests -x -q -m gpu 2>&1 | tail -1
python tools/profile_target.py --sweeps 10 --lod 3 --time 2>&1 | tail -1
SLK_LOD_CTA_THREADS=512 python tools/profile_target.py --sweeps 0 --lod 3 --time 2>&1 | tail -1
SLK_LOD_CTA_THREADS=768 SLK_LOD_SMEM_DOUBLES=4624 python tools/profile_target.py --sweeps 0 --lod 3 --time 2>&1 | tail -1
