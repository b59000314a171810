"""A/B: chunked history runs with the launches on one stream vs alternating between two (RTS-79, 1e7 years, history every 10 years)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import powersystemsreliabilityassessment_b200 as P
from powersystemsreliabilityassessment_b200 import rts79
cap, mttf, mttr = rts79.units()
gens = [P.Generator(i + 1, float(c), float(a), float(b)) for i, (c, a, b) in enumerate(zip(cap, mttf, mttr))]
lm = P.LoadModel(rts79.load_curve_int().astype(np.float64))
Y = 10_000_000
for single in (True, False, True, False):
    with P.Engine(single_stream=single) as e:
        P.run_sequential_mc(gens, lm, Y, seed=1, engine=e)
        ts = []
        for rep in range(6):
            t0 = time.perf_counter(); res, r = P.run_sequential_mc(gens, lm, Y, seed=1, year0=rep * Y, engine=e, details=True); ts.append(time.perf_counter() - t0)
        ref = e.seq_mc(Y, seed=1, year0=5 * Y, history=10)
        assert np.array_equal(ref.history, res.convergence_history)
        print(f"single_stream={single}: run_sequential_mc min {1e3 * min(ts):.2f} ms median {1e3 * sorted(ts)[3]:.2f} ms  kernel {r.kernel_ms:.2f} ms  -> {Y / min(ts) / 1e6:.1f} M years/s", flush=True)
