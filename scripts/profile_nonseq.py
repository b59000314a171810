"""Short driver for ncu: launches of the non-sequential kernel on RTS-79 (1e8 samples each)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from powersystemsreliabilityassessment_b200 import Engine, rts79
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
cap, mttf, mttr = rts79.units()
with Engine() as e:
    e.set_system(cap, mttf, mttr); e.set_load(rts79.load_curve_int())
    for i in range(3):
        r = e.nonseq_mc(n, seed=10 + i)
        print(i, r["kernel_ms"], n / r["kernel_ms"] * 1e3, r["lole"])
