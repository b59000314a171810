"""Short driver for ncu: a few launches of the > 32-unit sequential kernel on BASELINE config 5 (1024 units), or on k copies of
RTS-79 (usage: profile_wide.py <years> [k])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from powersystemsreliabilityassessment_b200 import Engine, rts79
years = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
c5 = rts79.synthetic_system(32, 37.0) if k == 32 else rts79.synthetic_system(k, 1.0 * k * 1.12)
with Engine() as e:
    e.set_system(c5[0], c5[1], c5[2]); e.set_load(c5[3])
    for i in range(3):
        r = e.seq_mc(years, seed=10 + i)
        print(i, r.kernel_ms, years / r.kernel_ms * 1e3, r.lole)
