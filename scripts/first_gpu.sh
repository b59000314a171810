set -x
make -C oracle -s
timeout 900 python -m pytest tests/test_gpu_seq.py -x -q 2>&1 | tail -30
SEGS=8736,4384,2944,1760 WPBS=16,24 timeout 900 python scripts/sweep_seq.py 4e6 2>&1 | tail -40
