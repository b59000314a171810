"""seq_wide.cu on mid-size systems (64 .. 256 units): warps per block / blocks per SM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from powersystemsreliabilityassessment_b200 import Engine, rts79
import sys as _s
for U in ((384, 512, 768, 1024) if len(_s.argv) > 1 else (64, 96, 128, 192, 256)):
    k = U // 32
    s = rts79.synthetic_system(k, 1.0 * k * 1.12)
    for wpb in ((3, 4, 5) if len(_s.argv) > 1 else (0, 1, 2, 3, 4, 6)):
        with Engine(warps_per_block=wpb) as e:
            e.set_system(s[0], s[1], s[2]); e.set_load(s[3])
            e.seq_mc(10_000, seed=1)
            r = e.seq_mc(1_000_000 if U <= 256 else 400_000, seed=42); r.kernel_ms *= (1.0 if U <= 256 else 2.5)
            print(f"{U:4d} units  warps/block {wpb}: {r.kernel_ms:8.2f} ms  {1e6 / r.kernel_ms * 1e3 / 1e6:7.2f} M yr/s", flush=True)
