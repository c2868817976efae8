echo "all phase 1"; python tools/ms_kernel_time.py 2>&1 | grep -E "kernel|cycles"
echo "all phase 0"; SLK_MS_DEBUG_PREV0=1 python tools/ms_kernel_time.py 2>&1 | grep -E "kernel|cycles"
echo "all phase 0, no PDL"; SLK_NO_PDL=1 SLK_MS_DEBUG_PREV0=1 python tools/ms_kernel_time.py 2>&1 | grep -E "kernel|cycles"
