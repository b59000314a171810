"""Round-2 features of the sequential path, every one through the C ABI against the CPU oracle:
static / queue phases of seq_wide.cu, non-fatal event-list overflow of seq_fast.cu, the in-kernel
ENS histogram and the one-launch VaR / CVaR from its counts (tail_risk.jl:168-175, seqMain.m:287; SURVEY a-12),
histogram exchange for cross-process sharding, the history launch plan at its chunk boundary, the lossless
Float64-load rule (PSA.jl:192,253) and the self-contained multi-area call."""
import numpy as np
import pytest

import powersystemsreliabilityassessment_b200 as P
from oracle import oracle as O
from powersystemsreliabilityassessment_b200 import DISC_MATLAB, Engine, rts79

pytestmark = pytest.mark.gpu

SUMS = ("sum_lol_hours", "sum_ens_fp", "sum_entries", "sum_lol_sq", "sum_ens_sq", "years_with_loss", "events")


def _same(a, b):
    assert np.array_equal(a.lol_hours, b.lol_hours) and np.array_equal(a.entries, b.entries)
    assert np.array_equal(a.raw["ens_fp_vector"], b.raw["ens_fp_vector"])
    for k in SUMS:
        assert a.raw[k] == b.raw[k], k


def test_wide_static_and_queue_phases_give_the_oracle_integers(engine):
    """seq_wide.cu generates the first blocks of every unit in a static lane = unit phase and the rest from per-warp
    to-do queues.  The split (psra_config.reserved[3]: from "one static block" to "almost everything static") and the
    block size must not change a single integer: oracle's literal loop, with and without the MATLAB discretisation,
    ragged hour count, unit counts that are not multiples of 32 (padding lanes) and config 5."""
    rng = np.random.default_rng(5)
    cases = []
    c5 = rts79.synthetic_system(32, 37.0)
    cases.append((c5[0], c5[1], c5[2], c5[3][:8736 - 45], 6))
    for U in (33, 70, 200):
        k = (U + 31) // 32
        s = rts79.synthetic_system(k, 1.05 * U / 32.0)
        pick = np.sort(rng.choice(32 * k, U, replace=False))
        cases.append((s[0][pick], s[1][pick] * rng.uniform(0.5, 1.5, U), s[2][pick], s[3], 12))
    for cap, mttf, mttr, load, ny in cases:
        engine.set_system(cap, mttf, mttr); engine.set_load(load)
        for disc in (0, DISC_MATLAB):
            ref = engine.seq_mc(ny * 8, seed=5, year0=40, init_mode=(0 if disc else 1) | disc, per_year=True, fail_count=True, group=10)
            assert ref.lol_hours.sum() > 0
            if disc:
                lol, ens, ent = O.seq_matlab_philox(cap, mttf, mttr, load.astype(np.float64), 5, 40, ny)
            else:
                lol, ens, ent = O.seq_philox(cap, mttf, mttr, load.astype(np.float64), 5, 40, ny, 1, 1)
            assert np.array_equal(ref.lol_hours[:ny].astype(np.float64), lol) and np.array_equal(ref.entries[:ny].astype(np.float64), ent)
            assert np.array_equal(ref.raw["ens_fp_vector"][:ny].astype(np.float64), ens)
            for kw in (dict(static_blocks=1), dict(static_blocks=56), dict(static_blocks=8, warps_per_block=1),
                       dict(warps_per_block=3), dict(warps_per_block=2, blocks_per_sm=1), dict(force_team=True)):
                with Engine(**kw) as v:
                    v.set_system(cap, mttf, mttr); v.set_load(load)
                    b = v.seq_mc(ny * 8, seed=5, year0=40, init_mode=(0 if disc else 1) | disc, per_year=True, fail_count=True, group=10)
                _same(ref, b)
                assert np.array_equal(ref.fail_count, b.fail_count) and np.array_equal(ref.group_lol, b.group_lol)


def test_fast_kernel_event_list_overflow_is_replayed_not_fatal(rts):
    """psra_config.ev_cap far below the expected transitions of an RTS-79 year: most years overflow the warp's event
    list.  Single-segment kernel: those years go to the redo list; ring kernel (multi-year chains): the call is repeated
    with the generic kernel.  Either way the integers are the oracle's."""
    cap, mttf, mttr, load = rts["cap"], rts["mttf"], rts["mttr"], rts["load_int"]
    lol, ens, ent = O.seq_philox(cap, mttf, mttr, load.astype(np.float64), 21, 8, 48, 1, 1)
    with Engine(ev_cap=704) as small, Engine() as ref:
        for e in (small, ref):
            e.set_system(cap, mttf, mttr); e.set_load(load)
        a = small.seq_mc(3000, seed=21, year0=8, per_year=True, fail_count=True, group=10, history=10, tail_hist=True)
        b = ref.seq_mc(3000, seed=21, year0=8, per_year=True, fail_count=True, group=10, history=10, tail_hist=True)
        assert a.redone > 0 and b.redone == 0
        _same(a, b)
        assert np.array_equal(a.fail_count, b.fail_count) and np.array_equal(a.group_lol, b.group_lol) and np.array_equal(a.history, b.history)
        assert np.array_equal(a.lol_hours[:48].astype(np.float64), lol) and np.array_equal(a.raw["ens_fp_vector"][:48].astype(np.float64), ens)
        assert small.tail(None) == ref.tail(None)
        # ring variant: three-year chains
        l3, e3, n3 = O.seq_philox(cap, mttf, mttr, load.astype(np.float64), 21, 0, 8, 3, 0)
        with Engine(ev_cap=448) as tiny:       # the ring keeps one list per year-long segment: below the ~ 500 events of a year
            tiny.set_system(cap, mttf, mttr); tiny.set_load(load)
            c = tiny.seq_mc(24, seed=21, init_mode=0, years_per_chain=3, per_year=True)
        assert c.redone > 0
        assert np.array_equal(c.lol_hours.astype(np.float64), l3) and np.array_equal(c.raw["ens_fp_vector"].astype(np.float64), e3)
        assert np.array_equal(c.entries.astype(np.float64), n3)


@pytest.mark.parametrize("system", ["rts79", "c5", "generic"])
def test_tail_from_in_kernel_histogram_is_exact(engine, rts, system):
    """VaR / CVaR from the ENS histogram the MC kernel fills (no per-year vector) == the order statistics of the
    per-year vector (oracle numpy restatement of the a-12 spec), for the three kernel families."""
    if system == "rts79":
        engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(rts["load_int"])
        n, eng = 200_000, engine
    elif system == "c5":
        cap, mttf, mttr, load = rts79.synthetic_system(32, 37.0)
        engine.set_system(cap, mttf, mttr); engine.set_load(load)
        n, eng = 6_000, engine
    else:
        eng = Engine(force_generic=True)
        eng.set_system(rts["cap"], rts["mttf"], rts["mttr"]); eng.set_load(rts["load_int"])
        n = 20_000
    try:
        alphas = (0.0, 0.3, 0.5, 0.95, 0.99, 0.999, 1.0)
        r = eng.seq_mc(n, seed=31, per_year=True, tail_hist=True)
        res, hist = eng.tail(None, alphas=alphas, n_bins=50, bin_width=1000)
        x = r.raw["ens_fp_vector"]
        for t in res:
            v, c = O.cvar(x, t["alpha"])
            assert t["var"] == v and abs(t["cvar"] - c) <= 1e-12 * max(c, 1.0)
            assert t["n_tail"] == int((x >= v).sum())
        assert np.array_equal(hist, np.bincount(np.minimum(x // 1000, 49), minlength=50)) and hist.sum() == n
        # the same run without any per-year output: identical tail
        eng.seq_mc(n, seed=31, tail_hist=True)
        assert eng.tail(None, alphas=alphas) == res
    finally:
        if eng is not engine:
            eng.close()


def test_histogram_range_overflow_is_reported(rts):
    """A histogram too short for the requested quantile: PSRA_E_OVERFLOW, never a wrong number; quantiles that lie
    inside the range stay exact (the years beyond it are carried as a count and a sum)."""
    with Engine(tail_bins=4096) as eng:
        eng.set_system(rts["cap"], rts["mttf"], rts["mttr"]); eng.set_load(rts["load_int"])
        r = eng.seq_mc(100_000, seed=4, per_year=True, tail_hist=True)
        x = r.raw["ens_fp_vector"]
        assert (x >= 4096).sum() > 100
        res = eng.tail(None, alphas=(0.5, 0.8))
        for t in res:
            v, c = O.cvar(x, t["alpha"])
            assert t["var"] == v and abs(t["cvar"] - c) <= 1e-12 * c
        with pytest.raises(P.PsraError) as ei:
            eng.tail(None, alphas=(0.99,))
        assert ei.value.code == -3


def test_histograms_of_shards_add_up(engine, rts):
    """Cross-process sharding exchanges O(bins) integers: export the shards' histograms, add them, import, psra_tail --
    equal to the single run over all years."""
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(rts["load_int"])
    engine.seq_mc(90_000, seed=8, tail_hist=True)
    whole = engine.tail(None, alphas=(0.95, 0.99))
    parts = []
    for y0, n in ((0, 30_000), (30_000, 25_000), (55_000, 35_000)):
        engine.seq_mc(n, seed=8, year0=y0, tail_hist=True)
        parts.append(engine.tail_hist_export())
    width = max(len(c) for c, _ in parts)
    tot = np.zeros(width, dtype=np.int64)
    for c, _ in parts:
        tot[:len(c)] += c
    meta = sum(m for _, m in parts)
    assert meta[0] == 90_000 and width < 200_000
    engine.tail_hist_import(tot, meta)
    assert engine.tail(None, alphas=(0.95, 0.99)) == whole


def test_history_launch_plan_at_the_chunk_boundary(engine, rts):
    """Years that leave a remainder behind the last full history group used to produce a ninth launch
    (PSRA_MAX_CHUNKS = 8).  The history of such a run equals the running mean of its group sums."""
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(rts["load_int"])
    for n in (12_124_165, 8 * 21_504 * 10 * 8 + 5, (1 << 20) + 7):
        r = engine.seq_mc(n, seed=2, history=10, group=10)
        ref = np.cumsum(r.group_lol[:n // 10]) / (10.0 * np.arange(1, n // 10 + 1))
        assert r.history.shape == (n // 10,) and np.array_equal(r.history, ref)
        assert r.raw["sum_lol_hours"] == int(r.group_lol.sum())


def test_float_load_keeps_the_reference_loss_test(rts):
    """PSA.jl:192,253 compare Float64 capacity and load.  The drop-in entry points put a fractional load on the grid
    with ceil (c < L <=> c < ceil(L) for whole-grid capacities): the RTS-79 curve in MW (1530.76977 at hour 1) gives the
    float-load analytical LOLE 9.3941 h/yr, not the 9.3677 of the rint-ed curve."""
    gens = [P.Generator(i + 1, float(c), float(a), float(b)) for i, (c, a, b) in enumerate(zip(rts["cap"], rts["mttf"], rts["mttr"]))]
    lm = P.LoadModel(rts["load_mw"])
    with Engine() as eng:
        res, idx = P.run_sequential_mc(gens, lm, 10_000_000, seed=5, engine=eng, details=True)
        assert abs(res.lole_hours_yr - 9.3941103566) < 4 * idx.lole_se and idx.lole_se < 0.006
        ns = P.run_non_sequential_mc(gens, lm, 200_000_000, seed=5, engine=eng)
        ana = P.run_analytical(gens, lm, step_size=1.0, engine=eng)
        assert abs(ana.lole_hours_yr - 9.3941103566) < 1e-8
        assert abs(ns.lole_hours_yr - ana.lole_hours_yr) < 0.03
        # per year, the loss hours are those of the literal Float64 loop on the same streams
        r = eng.seq_mc(24, seed=6, per_year=True)
        lol, _, ent = O.seq_philox(rts["cap"], rts["mttf"], rts["mttr"], rts["load_mw"], 6, 0, 24, 1, 1)
        assert np.array_equal(r.lol_hours.astype(np.float64), lol) and np.array_equal(r.entries.astype(np.float64), ent)


def test_multi_area_call_leaves_the_engine_alone(rts):
    """psra_multi_area_mc used to replace the handle's unit table (stale sizes on the host side -> heap overflow in a
    later states read-back).  It now works on private tables."""
    with Engine() as eng:
        eng.set_system(rts["cap"], rts["mttf"], rts["mttr"]); eng.set_load(rts["load_int"])
        before = eng.seq_mc(2000, seed=9, per_year=True)
        ns_before = eng.nonseq_mc(4096, seed=9, per_sample=True, states=True)
        rng = np.random.default_rng(1)
        U = 70
        ua = rng.integers(0, 3, U); cap = rng.integers(20, 200, U).astype(float)
        loads = rng.integers(500, 2500, (3, 1000)).astype(float)
        topo = np.array([[0, 100, 50], [100, 0, 0], [50, 0, 0]], dtype=float)
        m = eng.multi_area_mc(ua, cap, rng.uniform(500, 2000, U), rng.uniform(20, 80, U), loads, topo, P.INTERCONNECTED, 64, seed=1)
        assert m["sum_lol_hours"].sum() > 0
        after = eng.seq_mc(2000, seed=9, per_year=True)
        _same(before, after)
        ns_after = eng.nonseq_mc(4096, seed=9, per_sample=True, states=True)
        assert np.array_equal(ns_before["states"], ns_after["states"]) and np.array_equal(ns_before["lol_hours"], ns_after["lol_hours"])
        imp, cnt, _ = eng.seq_unit_importance(500, seed=9)
        assert cnt.shape == (32,)
