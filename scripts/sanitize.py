"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import powersystemsreliabilityassessment_b200 as P
from powersystemsreliabilityassessment_b200 import rts79, DISC_MATLAB, INIT_ALL_UP
cap, mttf, mttr = rts79.units(); load = rts79.load_curve_int()
with P.Engine() as e:
    e.set_system(cap, mttf, mttr); e.set_load(load)
    print("seq fast", e.seq_mc(4000, seed=1, per_year=True, fail_count=True, group=10).lole)
    print("seq fast chain", e.seq_mc(4000, seed=1, years_per_chain=8, init_mode=0).lole)
    print("seq fast matlab", e.seq_mc(2000, seed=1, init_mode=INIT_ALL_UP | DISC_MATLAB).lole)
    rng = np.random.default_rng(0)
    d = np.maximum(rng.exponential(1.0, (8, 32, 200)) * np.where(np.arange(200) % 2 == 0, mttf[None, :, None], mttr[None, :, None]), 1e-9)
    print("seq injected", e.seq_eval_injected(d, years_per_chain=2).lole)
    print("nonseq", e.nonseq_mc(200000, seed=2, per_sample=True, states=True, group=100)["lole"])
    lam = 1 / mttf; mu = 1 / mttr; q = lam / (lam + mu)
    pr = e.copt(cap, q, 1.0); print("copt", e.copt_indices(pr, 1.0, 3405.0, rts79.load_curve_mw()))
    print("fd", e.fd_recursion(cap, mttf + mttr, mttr)[0][556])
    print("markov", e.markov2(1000.0, 50.0)[-1], e.dtmc_capacity([1000.0, 800.0], [50.0, 40.0], [100.0, 50.0], rng.random((500, 2)))[-1])
    r = e.seq_mc(20000, seed=3, keep_on_device=True); print("tail", e.tail(None, n_bins=20, bin_width=2000)[0])
    c5 = rts79.synthetic_system(32, 37.0)
    e.set_system(c5[0], c5[1], c5[2]); e.set_load(c5[3])
    print("seq wide", e.seq_mc(300, seed=4, per_year=True, fail_count=True, group=10, history=10).lole, "seq team chain", e.seq_mc(120, seed=4, years_per_chain=4).lole)
    gens = [P.DetailedGenerator("A", 400.0, 0.02, 4), P.DetailedGenerator("B", 300.0, 0.04, 3), P.DetailedGenerator("H", 200.0, 0.01, 2, 1e4)]
    base = 500.0 + 150.0 * np.sin(np.arange(8760) / 8760 * 2 * math.pi)
    P.schedule_maintenance(gens, [base[(w - 1) * 168:w * 168].max() for w in range(1, 53)])
    print("detailed", e.detailed_mc(gens, base, 30.0, 512, seed=5)[0].mean())
    ua = np.array([0] * 5 + [1] * 5 + [2] * 3); acap = np.array([400.0] * 5 + [200.0] * 5 + [100.0] * 3)
    amttf = np.array([1000.0] * 5 + [900.0] * 5 + [300.0] * 3); amttr = np.array([50.0] * 5 + [60.0] * 5 + [30.0] * 3)
    x = np.linspace(0, 2 * np.pi, 8760)
    loads = np.stack([np.rint(1000 + 500 * np.sin(x)), np.rint(800 + 400 * np.sin(x)), np.rint(150 + 100 * np.cos(x))])
    topo = np.array([[0, 200, 50], [200, 0, 30], [50, 30, 0]], float)
    print("multi-area", [e.multi_area_mc(ua, acap, amttf, amttr, loads, topo, pol, 200, seed=6, per_year=True)["lole"] for pol in (0, 1)])
    print("failure times", len(e.failure_times(1e-3, 4000)))
    e.set_system(cap, mttf, mttr); e.set_load(load)
    print("seq fast long history", e.seq_mc(1_100_000, seed=1, history=10).history[-1])
with P.Engine(force_team=True) as t:
    t.set_system(c5[0], c5[1], c5[2]); t.set_load(c5[3])
    print("seq team", t.seq_mc(200, seed=4, per_year=True).lole)
with P.Engine(unpacked_words=True) as g:
    g.set_system(cap, mttf, mttr); g.set_load(load)
    print("seq fast unpacked", g.seq_mc(2000, seed=1).lole)
with P.Engine(force_generic=True) as g:
    g.set_system(cap, mttf, mttr); g.set_load(load)
    print("seq generic", g.seq_mc(2000, seed=1, years_per_chain=4).lole)
# round 2: in-kernel ENS histogram + one-launch tail, redo path of seq_fast.cu (small event lists), seq_wide.cu with a unit
# count that is not a multiple of 32 and several block sizes, the staged history read-back
with P.Engine(ev_cap=704) as s:
    s.set_system(cap, mttf, mttr); s.set_load(load)
    r = s.seq_mc(3000, seed=21, per_year=True, fail_count=True, group=10, history=10, tail_hist=True)
    print("seq fast redo", r.redone, s.tail(None, alphas=(0.5, 0.95))[0]["var"])
for kw in (dict(), dict(warps_per_block=2), dict(warps_per_block=6), dict(static_blocks=56)):
    with P.Engine(**kw) as w:
        k = 70
        w.set_system(np.tile(cap, 3)[:k], np.tile(mttf, 3)[:k], np.tile(mttr, 3)[:k]); w.set_load(np.rint(2.1 * rts79.load_curve_mw()).astype(np.int32)[:8000])
        r = w.seq_mc(200, seed=4, per_year=True, fail_count=True, group=10, history=10, tail_hist=True)
        print("seq wide 70 units", kw, r.lole, w.tail(None, alphas=(0.9,))[0]["var"])
print("SANITIZE_RUN_OK")
