"""Per-phase cycle shares of seq_wide.cu on config 5 (needs a build with EXTRA=-DWIDE_PROFILE):
static generation, queue generation, wait at barrier 1, word sums, wait at barrier 2, evaluation + clear, wait at barrier 3,
per-year epilogue.  usage: make -C powersystemsreliabilityassessment_b200/csrc clean all EXTRA=-DWIDE_PROFILE; python scripts/wide_phases.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from powersystemsreliabilityassessment_b200 import Engine, rts79
c5 = rts79.synthetic_system(32, 37.0)
names = ("static", "queue", "wait B1", "word sums", "wait B2", "eval+clear", "wait B3", "epilogue")
for kw in (dict(), dict(warps_per_block=6), dict(blocks_per_sm=1), dict(blocks_per_sm=2)):
    with Engine(**kw) as e:
        e.set_system(c5[0], c5[1], c5[2]); e.set_load(c5[3])
        e.seq_mc(20_000, seed=1)
        r = e.seq_mc(200_000, seed=42)
        out = (C.c_uint64 * 32)()
        e._check(e._L.psra_last_counters(e._h, out, 32))
        v = [out[16 + i] for i in range(8)]
        tot = sum(v) or 1
        print(kw, f"{200_000 / r.kernel_ms * 1e3 / 1e6:.2f} M yr/s", "  ".join(f"{n} {100 * x / tot:.1f}%" for n, x in zip(names, v)), flush=True)
