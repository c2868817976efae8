SLK_MS_TIMELINE=1 python tools/profile_target.py --msweeps 1 2>&1 | grep -E "pair|M-sweep" | head -24
