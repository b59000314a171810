"""End-to-end breakdown of the one-call multi-GPU path (Engine(ngpus=G)): set_generators, seq_mc without / with history."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import powersystemsreliabilityassessment_b200 as P
from powersystemsreliabilityassessment_b200 import rts79
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cap, mttf, mttr = rts79.units()
gens = [P.Generator(i + 1, float(c), float(a), float(b)) for i, (c, a, b) in enumerate(zip(cap, mttf, mttr))]
lm = P.LoadModel(rts79.load_curve_int().astype(np.float64))
Y = 10_000_000 * G
with P.Engine(ngpus=G) as e:
    for rep in range(4):
        t0 = time.perf_counter(); e.set_generators(gens, lm); t1 = time.perf_counter()
        r = e.seq_mc(Y, seed=1, year0=rep * Y); t2 = time.perf_counter()
        r2 = e.seq_mc(Y, seed=1, year0=rep * Y, history=10); t3 = time.perf_counter()
        res = P.run_sequential_mc(gens, lm, Y, seed=1, year0=rep * Y, engine=e); t4 = time.perf_counter()
        buf = np.empty(Y // 10); t5 = time.perf_counter(); buf[:] = 1.0; t6 = time.perf_counter()
        print(f"G={G} set_generators {1e3*(t1-t0):.2f} ms | seq_mc {1e3*(t2-t1):.2f} ms (kernel {r.kernel_ms:.2f}) | seq_mc+history {1e3*(t3-t2):.2f} ms "
              f"(kernel {r2.kernel_ms:.2f}) | run_sequential_mc {1e3*(t4-t3):.2f} ms | first touch of a fresh {Y // 10 * 8 >> 20} MB buffer {1e3*(t6-t5):.2f} ms", flush=True)
