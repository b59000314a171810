"""Occupancy experiment for seq_wide.cu: config 5 cut to a half year (4368 hours: the int32 timeline is 17.5 KB, so up to 8
blocks fit an SM), blocks per SM swept.  Build with EXTRA=-DWIDE_BPS4=8 for the 64-register variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from powersystemsreliabilityassessment_b200 import Engine, rts79
c5 = rts79.synthetic_system(32, 37.0)
for H in (4368, 8736):
    for bps in (2, 4, 5, 6, 7, 8):
        with Engine(blocks_per_sm=bps) as e:
            e.set_system(c5[0], c5[1], c5[2]); e.set_load(c5[3][:H])
            e.seq_mc(20_000, seed=1)
            r = e.seq_mc(400_000, seed=42)
            print(f"H {H} blocks/SM <= {bps}: {r.kernel_ms:8.2f} ms  {400_000 / r.kernel_ms * 1e3 / 1e6:7.3f} M yr/s  jobs/yr {e.last_counters()['jobs'] / 400_000:.0f}", flush=True)
