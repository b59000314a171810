"""seq_wide.cu on mid-size systems (k x RTS-79): years/s against the block size (warps per block)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from powersystemsreliabilityassessment_b200 import Engine, rts79
for k in (2, 3, 6, 10, 32):
    sysk = rts79.synthetic_system(k, 1.16 * k)
    years = int(4e6 / k)
    for wpb in (0, 1, 2, 3, 4, 6):
        with Engine(warps_per_block=wpb) as e:
            e.set_system(sysk[0], sysk[1], sysk[2]); e.set_load(sysk[3])
            e.seq_mc(2000, seed=1)
            best = min(e.seq_mc(years, seed=2 + i).kernel_ms for i in range(2))
            r = e.seq_mc(years, seed=2)
            print(f"units {32 * k:5d} warps/block {wpb}: {years / best * 1e3 / 1e6:7.2f} M yr/s  lole {r.lole:.3f}", flush=True)
    with Engine(force_team=True) as e:
        e.set_system(sysk[0], sysk[1], sysk[2]); e.set_load(sysk[3])
        e.seq_mc(2000, seed=1)
        best = min(e.seq_mc(years // 4, seed=2 + i).kernel_ms for i in range(2))
        print(f"units {32 * k:5d} seq_team: {years // 4 / best * 1e3 / 1e6:7.2f} M yr/s", flush=True)
