SLK_MS_TIMELINE=1 python tools/profile_target.py --msweeps 1 2>&1 | grep -E "slow CTA|SMs with|step kernel CTAs" | tail -14
