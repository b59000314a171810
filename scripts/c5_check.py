"""Config 5: synthetic 1024-unit system (RTS-79 x 32, load x 37): throughput of the generic kernel + statistics."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from powersystemsreliabilityassessment_b200 import Engine, rts79
years = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000
cap, mttf, mttr, load = rts79.synthetic_system(32, 37.0)
with Engine() as e:
    e.set_system(cap, mttf, mttr); e.set_load(load)
    e.seq_mc(2000, seed=1)
    t0 = time.time(); r = e.seq_mc(years, seed=2); dt = time.time() - t0
    print(f"C5 U=1024: {years} years kernel {r.kernel_ms:.1f} ms -> {years / r.kernel_ms * 1e3:.3e} years/s; LOLE {r.lole:.3f} +- {r.lole_se:.3f} "
          f"(analytical 8.033) EENS {r.eens:.1f} LOLF {r.lolf:.3f} events/yr {r.events / years:.1f}")
    lam = 1 / mttf; mu = 1 / mttr; q = lam / (lam + mu)
    t0 = time.time(); p = e.copt(cap, q, 1.0); l, eue = e.copt_indices(p, 1.0, float(cap.sum()), load.astype(float)); dt = time.time() - t0
    print(f"C5 COPT: {len(p)} states, LOLE {l:.6f} EUE {eue:.3f} in {dt * 1e3:.1f} ms")
    g = e.nonseq_mc(1_000_000, seed=3)
    print(f"C5 nonseq 1e6 samples: {g['kernel_ms']:.2f} ms, LOLE {g['lole']:.3f} +- {g['lole_se']:.3f}")
cap, mttf, mttr = rts79.units()
with Engine() as e:
    e.set_system(cap, mttf, mttr); e.set_load(rts79.load_curve_int())
    for n in (100_000, 10_000_000, 100_000_000):
        g = e.nonseq_mc(n, seed=5)
        print(f"RTS nonseq {n}: {g['kernel_ms']:.3f} ms -> {n / g['kernel_ms'] * 1e3:.3e} samples/s LOLE {g['lole']:.4f} +- {g['lole_se']:.4f}")
    r = e.seq_mc(1_000_000, seed=9, per_year=False, keep_on_device=True)
    t0 = time.time(); res = e.tail(None, alphas=(0.95, 0.99)); dt = time.time() - t0
    print("tail 1e6 years:", res, f"{dt * 1e3:.1f} ms")
    # bias check: 10 seeds x 1e8 years
    tot = None
    for s in range(10):
        r = e.seq_mc(100_000_000, seed=1000 + s)
        tot = r.raw if tot is None else {k: tot[k] + r.raw[k] for k in tot}
    from powersystemsreliabilityassessment_b200 import indices_from_raw
    idx = indices_from_raw(tot)
    print(f"1e9 years: LOLE {idx.lole:.5f} +- {idx.lole_se:.5f} (analytical 9.36774) EENS {idx.eens:.3f} +- {idx.eens_se:.3f} (1176.181) LOLF {idx.lolf:.5f} LOLD {idx.lold:.4f}")
