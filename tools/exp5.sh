set -e
run() { python tools/profile_target.py --sweeps 10 --lod 0 --time 2>&1 | tail -2 | grep -o "ls_blocks_per_sm': [0-9.]*\|ls_cta_smem': [0-9.]*\|sweep ms [0-9.]*\|Error.*" | tr '\n' ' '; echo; }
echo "== NS=2 maxthreads=768"; SLK_NVCC_DEFS="-DSLK_TILE_NS=2 -DSLK_LS_MAXTHREADS=768" python -m swiftlink_b200.build --force > /dev/null
echo -n "cta768: "; SLK_LS_CTA_THREADS=768 run
echo -n "cta768 team128 (6 teams): "; SLK_LS_CTA_THREADS=768 SLK_LS_TEAM=128 run
echo -n "cta768 smem 2988: "; SLK_LS_CTA_THREADS=768 SLK_LS_SMEM_DOUBLES=2988 run
echo "== NS=1 maxthreads=1024"; SLK_NVCC_DEFS="-DSLK_TILE_NS=1 -DSLK_LS_MAXTHREADS=1024" python -m swiftlink_b200.build --force > /dev/null
echo -n "cta1024 team256 (4 teams) smem 2988: "; SLK_LS_CTA_THREADS=1024 SLK_LS_SMEM_DOUBLES=2988 run
echo -n "cta768 team256: "; SLK_LS_CTA_THREADS=768 run
python -m swiftlink_b200.build --force > /dev/null
