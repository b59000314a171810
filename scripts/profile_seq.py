"""Short driver for ncu: a few launches of the sequential kernel on RTS-79 (1e6 years each)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from powersystemsreliabilityassessment_b200 import Engine, rts79
years = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
cap, mttf, mttr = rts79.units()
with Engine() as e:
    e.set_system(cap, mttf, mttr); e.set_load(rts79.load_curve_int())
    for i in range(4):
        r = e.seq_mc(years, seed=10 + i)
        print(i, r.kernel_ms, years / r.kernel_ms * 1e3, r.lole)
