set -x
timeout 600 python -m pytest tests/test_gpu_seq.py -x -q -k "team or sharding" 2>&1 | tail -5
timeout 300 python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from powersystemsreliabilityassessment_b200 import Engine, rts79
c5 = rts79.synthetic_system(32, 37.0)
for kw in (dict(), dict(force_team=True)):
    with Engine(**kw) as e:
        e.set_system(c5[0], c5[1], c5[2]); e.set_load(c5[3])
        e.seq_mc(2000, seed=1)
        for n in (200_000, 1_000_000):
            r = e.seq_mc(n, seed=42)
            print(kw, n, round(r.kernel_ms, 2), "ms", round(n / r.kernel_ms * 1e3 / 1e6, 3), "M yr/s", r.lole, e.last_counters())
PY
