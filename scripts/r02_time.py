"""Kernel timings of the sequential kernels and their variants (round 2): config 5 (seq_wide.cu packed / int32 timeline,
blocks per SM), RTS-79 (seq_fast.cu), histogram overhead, tail wall time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from powersystemsreliabilityassessment_b200 import Engine, rts79

c5 = rts79.synthetic_system(32, 37.0)
variants = [("wide default", dict())] + [(f"wide theta {(k - 16) / 8:+.3f}", dict(static_blocks=k)) for k in (8, 12, 16, 18, 24, 32, 48)] + \
           [(f"wide {w} warps", dict(warps_per_block=w)) for w in (2, 3, 5, 6, 7, 8)] + [("wide bps4", dict(blocks_per_sm=4)), ("wide 6 warps bps3", dict(warps_per_block=6, blocks_per_sm=3))]
if len(sys.argv) > 1 and sys.argv[1] == "short":
    variants = variants[:1]
if len(sys.argv) > 1 and sys.argv[1] == "warps":
    variants = [v for v in variants if "warps" in v[0] or "default" in v[0] or "bps" in v[0]]
for name, kw in variants:
    with Engine(**kw) as e:
        e.set_system(c5[0], c5[1], c5[2]); e.set_load(c5[3])
        e.seq_mc(20_000, seed=1)
        for n in (200_000, 1_000_000):
            r = e.seq_mc(n, seed=42)
            print(f"{name:24s} {n:8d} yr  {r.kernel_ms:9.2f} ms  {n / r.kernel_ms * 1e3 / 1e6:7.3f} M yr/s  LOLE {r.lole:.4f}  {e.last_counters()}", flush=True)

cap, mttf, mttr = rts79.units()
load = rts79.load_curve_int()
with Engine() as e:
    e.set_system(cap, mttf, mttr); e.set_load(load)
    e.seq_mc(100_000, seed=1)
    for n in (1_000_000, 10_000_000):
        r = e.seq_mc(n, seed=42)
        print(f"fast                     {n:8d} yr  {r.kernel_ms:9.2f} ms  {n / r.kernel_ms * 1e3 / 1e6:7.2f} M yr/s  LOLE {r.lole:.4f}  {e.last_counters()}", flush=True)
        r = e.seq_mc(n, seed=42, tail_hist=True)
        t0 = time.perf_counter(); t = e.tail(None); dt = time.perf_counter() - t0
        print(f"fast + ENS histogram     {n:8d} yr  {r.kernel_ms:9.2f} ms  tail wall {dt * 1e3:.3f} ms  VaR95 {t[0]['var']:.0f} CVaR95 {t[0]['cvar']:.1f} VaR99 {t[1]['var']:.0f} CVaR99 {t[1]['cvar']:.1f}", flush=True)
    for U in (64, 96, 128):
        k = U // 32
        s = rts79.synthetic_system(k, 1.0 * k * 1.12)
        e.set_system(s[0], s[1], s[2]); e.set_load(s[3])
        e.seq_mc(10_000, seed=1)
        r = e.seq_mc(1_000_000, seed=42)
        print(f"{U} units                 1000000 yr  {r.kernel_ms:9.2f} ms  {1e6 / r.kernel_ms * 1e3 / 1e6:7.2f} M yr/s  LOLE {r.lole:.4f} redone {r.redone}", flush=True)
