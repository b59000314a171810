set -x
nvidia-smi -L
make -C oracle -s
timeout 900 python -m pytest tests/test_gpu_seq.py -x -q 2>&1 | tail -30
timeout 600 python scripts/sweep_seq.py 2e6 2>&1 | tail -30
