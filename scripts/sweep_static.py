"""Sweep of the statically scheduled blocks per unit (single-segment seq_fast kernel) on RTS-79."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from powersystemsreliabilityassessment_b200 import Engine, rts79
cap, mttf, mttr = rts79.units()
for sb in (0, 1, 2, 3, 4, 5):
    with Engine(static_blocks=sb) as e:
        e.set_system(cap, mttf, mttr); e.set_load(rts79.load_curve_int())
        e.seq_mc(100000, seed=1)
        best = min(e.seq_mc(10_000_000, seed=10 + i).kernel_ms for i in range(3))
        c = e.last_counters()
        print(sb, round(best, 2), "ms", round(1e7 / best / 1e3, 1), "M yr/s waves", c["waves"] / 1e7, "jobs", c["jobs"] / 1e7, "max slots", c["pend_max"])
