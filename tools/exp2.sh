python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for cfg in "0 512 512" "6060 256 256" "6060 128 128" "6060 512 512"; do
  set -- $cfg
  echo "== smem_doubles=$1 team=$2 cta=$3"
  if [ "$1" = "0" ]; then SLK_LS_TEAM=$2 SLK_CTA_THREADS=$3 python tools/profile_target.py --sweeps 10 --lod 1 --time 2>&1 | tail -1
  else SLK_LS_SMEM_DOUBLES=$1 SLK_LS_TEAM=$2 SLK_CTA_THREADS=$3 python tools/profile_target.py --sweeps 10 --lod 0 --time 2>&1 | tail -1; fi
done
