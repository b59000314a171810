"""Regenerates tests/golden/*.  The reference is Julia/MATLAB and cannot be executed in this image,
so the fixtures are outputs of the pinned CPU oracle (oracle/psra_oracle.c, itself checked against the
SURVEY.md 8c known answers in tests/test_oracle.py); they freeze the oracle against regressions and
travel to the GPU box, where /root/reference does not exist."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import injected_durations
from oracle import oracle as O
from powersystemsreliabilityassessment_b200 import rts79

cap, mttf, mttr = rts79.units()
lam = 1 / mttf; mu = 1 / mttr; q = lam / (lam + mu)
load = rts79.load_curve_int().astype(float)
os.makedirs(os.path.join(ROOT, "tests/golden"), exist_ok=True)
np.save(os.path.join(ROOT, "tests/golden/rts79_copt_step10.npy"), O.copt_build(cap, q, 10.0))
rng = np.random.default_rng(123)
dur = injected_durations(rng, mttf, mttr, 1, 200)[0]
lol, eue, ent, _ = O.seq_literal(cap, load, 3, dur)
np.savez(os.path.join(ROOT, "tests/golden/seq_literal_seed123.npz"), lol=lol, eue=eue, ent=ent, dur=dur)
lol, eue, ent = O.seq_philox(cap, mttf, mttr, load, 42, 0, 64, 1, 1)
np.savez(os.path.join(ROOT, "tests/golden/seq_philox_seed42.npz"), lol=lol, eue=eue, ent=ent)
# sampler specification, draw by draw: edge draws + seeded random draws, two means
rng = np.random.default_rng(2026)
edge = [0, 1, 2, 3, 0xFFFFFFFF, 0xFFFFFFFE, 0x80000000, 0x7FFFFFFF, 0xB504F333, 0xB504F334, 0x5A827999, 0x00FFFFFF, 0x01000000]
edge += [1 << k for k in range(32)] + [(1 << k) - 1 for k in range(1, 32)]
draws = np.concatenate([np.array(edge, dtype=np.uint64), rng.integers(0, 1 << 32, 4000, dtype=np.uint64)]).astype(np.uint32)
t450, eb = O.sampler_durations(450.0, draws)
t2940, _ = O.sampler_durations(2940.0, draws)
np.savez(os.path.join(ROOT, "tests/golden/sampler_draws.npz"), draws=draws, e_bits=eb, ticks_450=t450, ticks_2940=t2940)
# non-sequential sampler (PSA.jl:169-208 on the Philox stream): 256 samples
nl, ne, st = O.nonseq_philox(cap, mttf, mttr, load, 7, 1000, 256)
np.savez(os.path.join(ROOT, "tests/golden/nonseq_philox_seed7.npz"), lol=nl, eue=ne, states=st)
print("golden fixtures written")
