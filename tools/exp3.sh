for cfg in "6060 256 256" "11180 512 512" "11180 256 256" "6060 256 512" "2988 256 256" "2988 128 128" "940 128 128" "940 256 256"; do
  set -- $cfg
  echo "== smem_doubles=$1 team=$2 cta=$3"
  SLK_LS_SMEM_DOUBLES=$1 SLK_LS_TEAM=$2 SLK_CTA_THREADS=$3 python tools/profile_target.py --sweeps 10 --lod 0 --time 2>&1 | tail -2 | grep -o "ls_blocks_per_sm': [0-9.]*\|sweep ms [0-9.]*" | tr '\n' ' '; echo
done
