"""Kernel time of the detailed MC (tail_risk.jl's 6-unit system, as bench.py's configs.f1_detailed_mc) at 1e5 and 2e6 years."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import powersystemsreliabilityassessment_b200 as P
dg = [P.DetailedGenerator("Nuclear", 400.0, 0.02, 4), P.DetailedGenerator("Coal_A", 300.0, 0.04, 3),
      P.DetailedGenerator("Coal_B", 300.0, 0.04, 3), P.DetailedGenerator("Gas", 150.0, 0.05, 2),
      P.DetailedGenerator("Hydro_ELU", 200.0, 0.01, 2, 200.0 * 50.0), P.DetailedGenerator("Old_56", 56.0, 0.10, 0)]
rng = np.random.default_rng(7); hh = np.arange(1, 8761)
base = np.maximum(0.0, 750.0 + 300.0 * np.sin((hh - 2000) / 8760 * 2 * math.pi) + 50.0 * rng.standard_normal(8760))
P.schedule_maintenance(dg, [base[(w - 1) * 168:min(w * 168, 8760)].max() for w in range(1, 53)])
with P.Engine() as eng:
    eng.detailed_mc(dg, base, base.max() * 0.05, 20_000, seed=1)
    for n in (100_000, 2_000_000):
        yl, _, ms = eng.detailed_mc(dg, base, base.max() * 0.05, n, seed=1)
        print(f"{n:8d} years  {ms:8.2f} ms  {n / ms * 1e3 / 1e6:6.2f} M years/s  {n * 8760 / ms * 1e3 / 1e9:6.1f} G hour-steps/s  mean LOLE {yl.mean():.4f}", flush=True)
