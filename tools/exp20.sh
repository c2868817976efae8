python tools/ms_stress.py --markers 2000 --sweeps 9 --seed 77 2>&1 | grep -v WARN | tail -12
python tools/ms_stress.py --markers 10000 --sweeps 3 --seed 5 2>&1 | grep -v WARN | tail -6
