"""compute-sanitizer run over the entry points added in round 1i-1l (weak points, sampler diagnostic) and the kernels
touched there (seq_fast single-segment / ring variants, seq_wide, nonseq)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import powersystemsreliabilityassessment_b200 as P
from powersystemsreliabilityassessment_b200 import rts79
cap, mttf, mttr = rts79.units(); load = rts79.load_curve_int()
with P.Engine() as e:
    e.set_system(cap, mttf, mttr); e.set_load(load)
    print("importance", e.seq_unit_importance(600, seed=1, per_year=True)[0][21])
    print("importance chain", e.seq_unit_importance(400, seed=1, years_per_chain=4, init_mode=0)[0][21])
    print("sampler", e.sampler_durations(450.0, np.arange(0, 1 << 32, 1 << 20, dtype=np.uint64).astype(np.uint32))[0][:3])
    print("seq fast", e.seq_mc(3000, seed=1, per_year=True, fail_count=True, group=10).lole)
    print("seq fast chain", e.seq_mc(2000, seed=1, years_per_chain=8, init_mode=0).lole)
    print("nonseq", e.nonseq_mc(100000, seed=2)["lole"])
    c5 = rts79.synthetic_system(32, 37.0)
    e.set_system(c5[0], c5[1], c5[2]); e.set_load(c5[3])
    print("seq wide", e.seq_mc(200, seed=4, per_year=True).lole)
    c2 = rts79.synthetic_system(2, 2.3)
    e.set_system(c2[0], c2[1], c2[2]); e.set_load(c2[3])
    print("importance 64 units", e.seq_unit_importance(100, seed=4)[0].max(), "wide 64 units", e.seq_mc(300, seed=4).lole)
print("SANITIZE_RUN_OK")
