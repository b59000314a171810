"""ncu driver: config 5 cut to 4368 hours (8 blocks of seq_wide.cu per SM fit with the 64-register build -DWIDE_BPS4=8)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from powersystemsreliabilityassessment_b200 import Engine, rts79
c5 = rts79.synthetic_system(32, 37.0)
with Engine(blocks_per_sm=int(sys.argv[1]) if len(sys.argv) > 1 else 0) as e:
    e.set_system(c5[0], c5[1], c5[2]); e.set_load(c5[3][:4368])
    for i in range(3):
        r = e.seq_mc(200_000, seed=10 + i)
        print(i, r.kernel_ms, 200_000 / r.kernel_ms * 1e3)
