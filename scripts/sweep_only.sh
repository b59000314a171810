timeout 900 python scripts/sweep_seq.py 2e6 2>&1 | tail -40
