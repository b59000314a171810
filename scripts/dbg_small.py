import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from powersystemsreliabilityassessment_b200 import Engine
cap = np.array([10.0, 5.0, 7.0]); mttf = np.array([30.0, 20.0, 3.0]); mttr = np.array([10.0, 15.0, 2.0])
with Engine() as e:
    for H in (1, 31, 33, 100, 1000):
        load = np.full(H, 14, dtype=np.int32)
        e.set_system(cap, mttf, mttr); e.set_load(load)
        for ypc in (1, 4):
            print("H", H, "ypc", ypc, flush=True)
            r = e.seq_mc(32 * ypc, seed=H, init_mode=1, years_per_chain=ypc, per_year=True)
            print(r.lole, e.last_counters(), flush=True)
