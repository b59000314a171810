import sys, os
sys.path.insert(0, os.getcwd())
from powersystemsreliabilityassessment_b200 import Engine, rts79
cap, mttf, mttr = rts79.units()
with Engine() as e:
    e.set_system(cap, mttf, mttr); e.set_load(rts79.load_curve_int())
    for i in range(3):
        r = e.seq_mc(100_000_000, seed=100 + i)
        print(i, r.kernel_ms, e.last_counters())
