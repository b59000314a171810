make -C oracle -s
timeout 900 python -m pytest tests/test_gpu_seq.py -x -q 2>&1 | tail -15
timeout 600 python scripts/c5_check.py 200000 2>&1 | head -3
