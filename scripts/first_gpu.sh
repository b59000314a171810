set -x
make -C oracle -s
timeout 900 python -m pytest tests/test_gpu_seq.py -x -q 2>&1 | tail -30
timeout 900 python scripts/sweep_seq.py 4e6 2>&1 | tail -40
