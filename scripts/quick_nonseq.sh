set -x
timeout 600 python -m pytest tests/test_gpu_nonseq.py -x -q 2>&1 | tail -5
timeout 300 python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from powersystemsreliabilityassessment_b200 import Engine, rts79
cap, mttf, mttr = rts79.units()
for kw in (dict(), dict(force_generic=True)):
    with Engine(**kw) as e:
        e.set_system(cap, mttf, mttr); e.set_load(rts79.load_curve_int())
        e.nonseq_mc(1000, seed=1)
        for n in (100_000, 100_000_000, 1_000_000_000):
            r = e.nonseq_mc(n, seed=42)
            print(kw, n, round(r["kernel_ms"], 3), "ms", round(n / r["kernel_ms"] * 1e3 / 1e9, 2), "G samples/s", r["lole"])
PY
