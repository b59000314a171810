"""The five BASELINE.json configurations at their full sizes: bit-exact against the oracle's literal loops where the
oracle finishes in seconds (C1: 1e5 samples, C2: 1e4 years), and through size-independent properties where it does
not (C4: 1e6 years, C5: 1024 units, the bench's 1e7 years): shard concatenation / accumulator additivity,
accumulators == sums over the per-year vectors, exact order statistics of the tail reduction, confidence-interval
overlap with the analytical COPT values."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

SUMS = ("sum_lol_hours", "sum_ens_fp", "sum_entries", "sum_lol_sq", "sum_ens_sq", "years_with_loss", "events")


def _rts(engine, rts):
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(rts["load_int"])


def test_config1_nonsequential_1e5_samples_bit_exact(engine, rts):
    """C1: every one of the 1e5 samples (state word, LOL hours, ENS) equals the literal PSA.jl:169-208 loop."""
    _rts(engine, rts)
    n = 100_000
    g = engine.nonseq_mc(n, seed=42, per_sample=True, states=True)
    lol, ens, st = O.nonseq_philox(rts["cap"], rts["mttf"], rts["mttr"], rts["load_int"].astype(np.float64), 42, 0, n)
    assert np.array_equal(g["lol_hours"].astype(np.float64), lol)
    assert np.array_equal(g["ens"].astype(np.float64), ens)
    assert np.array_equal(g["states"].reshape(n, -1), st)
    assert g["lole"] == lol.sum() / n


def test_config2_sequential_1e4_years_bit_exact(engine, rts):
    """C2: all 1e4 years (LOL hours, ENS, deficit entries) equal the literal PSA.jl:214-269 hour/unit loop run on
    the same sampler streams; the indices follow from those integers."""
    _rts(engine, rts)
    n = 10_000
    r = engine.seq_mc(n, seed=42, per_year=True, history=10)
    lol, ens, ent = O.seq_philox(rts["cap"], rts["mttf"], rts["mttr"], rts["load_int"].astype(np.float64), 42, 0, n, 1, 1)
    assert np.array_equal(r.lol_hours.astype(np.float64), lol)
    assert np.array_equal(r.raw["ens_fp_vector"].astype(np.float64), ens)
    assert np.array_equal(r.entries.astype(np.float64), ent)
    assert r.lole == lol.sum() / n and r.eens == ens.sum() / n and r.lolf == ent.sum() / n
    # PSA.jl:263-265: running mean of the LOL hours every 10 years
    ref_hist = np.cumsum(lol)[9::10] / np.arange(10, n + 1, 10)
    assert np.allclose(r.history, ref_hist, rtol=1e-13, atol=0)


def test_config4_tail_1e6_years_properties(engine, rts):
    """C4: 1e6 years.  Accumulators == sums of the per-year vectors; two shards concatenate to the full run; VaR /
    CVaR are the exact order statistics of the vector; the histogram counts every year once."""
    _rts(engine, rts)
    n = 1_000_000
    r = engine.seq_mc(n, seed=2026, per_year=True, keep_on_device=True)
    res, hist = engine.tail(None, alphas=(0.95, 0.99), n_bins=64, bin_width=1000)
    x = r.raw["ens_fp_vector"]; lol = r.lol_hours.astype(np.int64); ent = r.entries.astype(np.int64)
    assert r.raw["sum_lol_hours"] == int(lol.sum()) and r.raw["sum_ens_fp"] == int(x.sum()) and r.raw["sum_entries"] == int(ent.sum())
    assert r.raw["sum_lol_sq"] == int((lol * lol).sum()) and r.raw["years_with_loss"] == int((lol > 0).sum())
    assert r.raw["sum_ens_sq"] == int((x.astype(object) ** 2).sum())
    assert np.array_equal(lol > 0, x > 0) and np.array_equal(lol > 0, ent > 0) and (ent <= lol).all()
    for t in res:
        v, c = O.cvar(x, t["alpha"])
        assert t["var"] == v and abs(t["cvar"] - c) <= 1e-12 * c and t["n_tail"] == int((x >= v).sum())
    assert hist.sum() == n and np.array_equal(hist, np.bincount(np.minimum(x // 1000, 63), minlength=64))
    a = engine.seq_mc(400_000, seed=2026, year0=0, per_year=True)
    b = engine.seq_mc(600_000, seed=2026, year0=400_000, per_year=True)
    assert np.array_equal(x, np.concatenate([a.raw["ens_fp_vector"], b.raw["ens_fp_vector"]]))
    assert np.array_equal(r.lol_hours, np.concatenate([a.lol_hours, b.lol_hours]))
    for k in SUMS:
        assert r.raw[k] == a.raw[k] + b.raw[k], k
    # statistics: analytical LOLE 9.3677 h/yr, EUE 1176.18 MWh/yr (integer load curve) within 4 standard errors
    assert abs(r.lole - 9.3677375218) < 4 * r.lole_se and abs(r.eens - 1176.181257) < 4 * r.eens_se


def test_bench_size_1e7_years_accumulators_are_additive(engine, rts):
    """The bench's step (1e7 RTS-79 years, accumulators only): any split into year ranges -- what the GPUs of a node
    each take -- adds up to the same integers, and the estimate agrees with the analytical value."""
    _rts(engine, rts)
    n = 10_000_000
    full = engine.seq_mc(n, seed=7)
    parts = [engine.seq_mc(c, seed=7, year0=y0) for y0, c in ((0, 1_250_000), (1_250_000, 3_750_001), (5_000_001, 4_999_999))]
    for k in SUMS:
        assert full.raw[k] == sum(p.raw[k] for p in parts), k
    assert abs(full.lole - 9.3677375218) < 4 * full.lole_se and abs(full.eens - 1176.181257) < 4 * full.eens_se
    # stationary alternating renewal processes: 2 H / (MTTF + MTTR) transitions per unit and year (462.4 in total)
    expected = float((2.0 * 8736.0 / (rts["mttf"] + rts["mttr"])).sum())
    assert abs(full.raw["events"] / n - expected) < 0.05


def test_config5_1024_units_properties(engine, rts):
    """C5: 32 x RTS-79 (1024 units), load scaled by 37 (analytical LOLE 8.033 h/yr): first years bit-exact vs the
    oracle, shard additivity at 2e5 years, CI overlap with the COPT value."""
    from powersystemsreliabilityassessment_b200 import rts79
    cap, mttf, mttr, load = rts79.synthetic_system(32, 37.0)
    engine.set_system(cap, mttf, mttr); engine.set_load(load)
    r0 = engine.seq_mc(12, seed=5, per_year=True)
    lol, ens, ent = O.seq_philox(cap, mttf, mttr, load.astype(np.float64), 5, 0, 12, 1, 1)
    assert np.array_equal(r0.lol_hours.astype(np.float64), lol) and np.array_equal(r0.raw["ens_fp_vector"].astype(np.float64), ens)
    assert np.array_equal(r0.entries.astype(np.float64), ent)
    n = 200_000
    full = engine.seq_mc(n, seed=5)
    a = engine.seq_mc(80_000, seed=5); b = engine.seq_mc(120_000, seed=5, year0=80_000)
    for k in SUMS:
        assert full.raw[k] == a.raw[k] + b.raw[k], k
    assert abs(full.lole - 8.033131075406098) < 4 * full.lole_se
    assert abs(full.eens - 13963.868303044579) < 4 * full.eens_se
    _rts(engine, rts)
