import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from powersystemsreliabilityassessment_b200 import Engine, rts79
cap, mttf, mttr, load = rts79.synthetic_system(32, 37.0)
with Engine() as e:
    e.set_system(cap, mttf, mttr); e.set_load(load)
    for seed in (3, 4, 5):
        g = e.nonseq_mc(20_000_000, seed=seed)
        print(seed, g['lole'], g['lole_se'], (g['lole'] - 8.033131) / g['lole_se'], g['eue'], g['kernel_ms'])
    r = e.seq_mc(2_000_000, seed=77)
    print("seq C5 2e6:", r.lole, r.lole_se, (r.lole - 8.033131) / r.lole_se, r.eens, "EUE analytical 13963.868", r.kernel_ms)
    print(e.last_counters())
