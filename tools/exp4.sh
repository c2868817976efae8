for cfg in "12816 512 512" "12816 256 256" "10000 256 256" "10000 256 512" "8000 256 256" "8000 128 128" "6000 256 256" "4000 128 128"; do
  set -- $cfg
  echo "== lod smem_doubles=$1 team=$2 cta=$3"
  SLK_LOD_SMEM_DOUBLES=$1 SLK_LOD_TEAM=$2 SLK_CTA_THREADS=$3 python tools/profile_target.py --sweeps 0 --lod 3 --time 2>&1 | tail -2 | grep -o "lod_blocks_per_sm': [0-9.]*\|lod_smem_doubles': [0-9.]*\|lod pass ms [0-9.]*" | tr '\n' ' '; echo
done
