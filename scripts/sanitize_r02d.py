"""compute-sanitizer run over the kernels touched in round 2d: nonseq_fast_kernel<kBlk> (template on the Philox block count; unit
counts 1..32), seq_wide.cu with the lane = hour resolution of flagged words (config 5, 64 / 96 units with and without the per-hour
failure counts, a system that loses load in most words), MATLAB discretisation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import powersystemsreliabilityassessment_b200 as P
from powersystemsreliabilityassessment_b200 import rts79, DISC_MATLAB, INIT_ALL_UP
cap, mttf, mttr = rts79.units(); load = rts79.load_curve_int()
with P.Engine() as e:
    for U in (1, 3, 4, 5, 13, 29, 32):
        e.set_system(cap[:U], mttf[:U], mttr[:U]); e.set_load(np.minimum(load, int(cap[:U].sum() * 0.8)))
        print("nonseq", U, e.nonseq_mc(20000, seed=2, per_sample=True, states=True)["lole"])
    c5 = rts79.synthetic_system(32, 37.0)
    e.set_system(c5[0], c5[1], c5[2]); e.set_load(c5[3])
    print("seq wide c5", e.seq_mc(300, seed=4, per_year=True, fail_count=True, group=10).lole)
    for k in (2, 3):
        s = rts79.synthetic_system(k, 1.0 * k * 1.12)
        e.set_system(s[0], s[1], s[2]); e.set_load(s[3])
        print("seq wide", 32 * k, e.seq_mc(400, seed=4, per_year=True, fail_count=True).lole, e.seq_mc(400, seed=4, tail_hist=True).lole)
        print("seq wide matlab", e.seq_mc(200, seed=4, init_mode=INIT_ALL_UP | DISC_MATLAB, per_year=True).lole)
    s = rts79.synthetic_system(2, 2.9)          # heavy load: loss of load in most words of the year
    e.set_system(s[0], s[1], s[2]); e.set_load(s[3][:1000])
    print("seq wide stressed, ragged hours", e.seq_mc(300, seed=9, per_year=True, fail_count=True).lole)
print("SANITIZE_RUN_OK")
