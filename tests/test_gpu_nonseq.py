"""GPU parity of the non-sequential state sampler against the oracle's literal PSA.jl:169-208 loop."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _for(mttf, mttr):
    lam = 1.0 / mttf; mu = 1.0 / mttr
    return lam / (lam + mu)


def test_injected_uniforms_bit_exact(engine, rts):
    rng = np.random.default_rng(0)
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(rts["load_int"])
    q = _for(rts["mttf"], rts["mttr"])
    r = rng.random((3000, 32))
    r[:200] *= 0.12            # force many outages so that loss samples are common
    r[0, :] = q                # boundary: rand() == FOR counts as UP (>=, PSA.jl:183)
    g = engine.nonseq_eval_uniforms(r, group=100)
    lol, eue, cp = O.nonseq_literal(rts["cap"], q, rts["load_int"].astype(float), r)
    assert np.array_equal(g["lol_hours"].astype(float), lol)
    assert np.array_equal(g["ens"].astype(float), eue)
    assert np.array_equal(g["cap"].astype(float), cp)
    assert g["cap"][0] == 3405
    assert g["raw"]["sum_lol_hours"] == int(lol.sum()) and g["raw"]["sum_ens_fp"] == int(eue.sum())
    assert g["raw"]["sum_ens_sq"] == sum(int(x) ** 2 for x in eue)
    assert np.array_equal(g["group_lol"], lol.reshape(-1, 100).sum(1).astype(np.int64))
    assert lol.sum() > 0


def test_injected_packed_states_bit_exact(engine, rts):
    rng = np.random.default_rng(1)
    cap = np.tile(rts["cap"], 3)[:70]; mttf = np.tile(rts["mttf"], 3)[:70]; mttr = np.tile(rts["mttr"], 3)[:70]
    load = np.rint(2.2 * rts["load_mw"]).astype(np.int32)
    engine.set_system(cap, mttf, mttr); engine.set_load(load)
    st = rng.integers(0, 2**32, size=(2000, 3), dtype=np.uint64).astype(np.uint32)
    st[:500] |= rng.integers(0, 2**32, size=(500, 3), dtype=np.uint64).astype(np.uint32)
    g = engine.nonseq_eval_states(st)
    lol, eue = O.nonseq_states(cap, load.astype(float), st)
    assert np.array_equal(g["lol_hours"].astype(float), lol)
    assert np.array_equal(g["ens"].astype(float), eue)
    assert (g["states"][:, 2] >> 6).max() == 0          # bits beyond U are masked off


def test_philox_bit_exact_and_sharding(engine, rts):
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(rts["load_int"])
    g = engine.nonseq_mc(5000, seed=77, sample0=1000, per_sample=True, states=True)
    lol, eue, st = O.nonseq_philox(rts["cap"], rts["mttf"], rts["mttr"], rts["load_int"].astype(float), 77, 1000, 5000)
    assert np.array_equal(g["lol_hours"].astype(float), lol)
    assert np.array_equal(g["ens"].astype(float), eue)
    assert np.array_equal(g["states"], st)
    a = engine.nonseq_mc(2000, seed=77, sample0=1000, per_sample=True)
    b = engine.nonseq_mc(3000, seed=77, sample0=3000, per_sample=True)
    assert np.array_equal(g["lol_hours"], np.concatenate([a["lol_hours"], b["lol_hours"]]))
    for k in ("sum_lol_hours", "sum_ens_fp", "sum_lol_sq", "sum_ens_sq", "samples_with_loss"):
        assert g["raw"][k] == a["raw"][k] + b["raw"][k]


def test_convergence_history_computed_on_device(engine, rts):
    """running mean of LOL hours every 100 iterations (PSA.jl:202-204)."""
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(rts["load_int"])
    for n in (100, 1234, 100_000):
        g = engine.nonseq_mc(n, seed=5, per_sample=True, history=100)
        k = n // 100
        want = np.cumsum(g["lol_hours"][: 100 * k].astype(np.float64)).reshape(k, 100)[:, -1] / (100.0 * np.arange(1, k + 1))
        assert g["history"].shape == (k,) and np.array_equal(g["history"], want)


def test_table_path_equals_generic_path(engine, rts):
    """<= 32 units: the table-driven kernel (byte tables + LOL-by-capacity table) and the generic one (masked sums +
    binary search) give the same integers, also for few units, a ragged unit count and the peak-load mode (H = 1)."""
    from powersystemsreliabilityassessment_b200 import Engine
    rng = np.random.default_rng(12)
    cases = [(rts["cap"], rts["mttf"], rts["mttr"], rts["load_int"]),
             (rts["cap"], rts["mttf"], rts["mttr"], np.array([2850], dtype=np.int32)),
             (rts["cap"][:5], rts["mttf"][:5] / 20, rts["mttr"][:5], rng.integers(0, 300, 100).astype(np.int32)),
             (rts["cap"][:27], rts["mttf"][:27] / 10, rts["mttr"][:27], rng.integers(1500, 3000, 777).astype(np.int32))]
    with Engine(force_generic=True) as gen:
        for cap, mttf, mttr, load in cases:
            engine.set_system(cap, mttf, mttr); engine.set_load(load)
            gen.set_system(cap, mttf, mttr); gen.set_load(load)
            a = engine.nonseq_mc(20_000, seed=4, sample0=77, per_sample=True, states=True, group=100)
            b = gen.nonseq_mc(20_000, seed=4, sample0=77, per_sample=True, states=True, group=100)
            for k in ("lol_hours", "ens", "states", "group_lol"):
                assert np.array_equal(a[k], b[k]), k
            assert a["raw"] == b["raw"] and a["lol_hours"].sum() > 0


def test_philox_many_units(engine, rts):
    from powersystemsreliabilityassessment_b200 import rts79
    cap, mttf, mttr, load = rts79.synthetic_system(32, 37.0)
    engine.set_system(cap, mttf, mttr); engine.set_load(load)
    g = engine.nonseq_mc(300, seed=3, per_sample=True, states=True)
    lol, eue, st = O.nonseq_philox(cap, mttf, mttr, load.astype(float), 3, 0, 300)
    assert np.array_equal(g["lol_hours"].astype(float), lol)
    assert np.array_equal(g["ens"].astype(float), eue)
    assert np.array_equal(g["states"], st)


def test_config1_statistics_hourly_and_peak(engine, rts):
    """BASELINE config 1: 1e5 samples vs the hourly curve and vs the annual peak (PLC*8760,
    Montecarlo_nsq_single/nsqMain.m:290-296); analytical targets from BASELINE.md section 3."""
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(rts["load_int"])
    g = engine.nonseq_mc(100_000, seed=42)
    assert abs(g["lole"] - 9.3677375218) < 3 * g["lole_se"]
    assert abs(g["eue"] - 1176.181257) < 3 * g["eue_se"]
    engine.set_load(np.array([2850], dtype=np.int32))
    p = engine.nonseq_mc(100_000, seed=42)
    plc = p["lole"]                       # H = 1: LOL "hours" per sample = loss indicator
    assert abs(plc - 0.084578060826) < 3 * p["lole_se"]
    assert 13.0 < p["eue"] < 16.0         # EDNS ballpark 14.51 MW


def test_adaptive_stop_rules(engine, rts):
    """f-2: CoV stop of seqMain.m:183-197 and beta stop of nsqMain.m:299-301 on top of batched launches."""
    import powersystemsreliabilityassessment_b200 as P
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(rts["load_int"])
    r = P.run_sequential_until_cov(engine, cov_threshold=0.05, batch_years=200, seed=3)
    assert 0 < r.cov_eens < 0.05 and r.years % 200 == 0 and 800 <= r.years <= 4000     # HL2 run stopped at 1245
    again = engine.seq_mc(r.years, seed=3)
    assert again.raw["sum_ens_fp"] == r.raw["sum_ens_fp"]        # batches are shards of one experiment
    engine.set_load(np.array([2850], dtype=np.int32))
    out, hist = P.run_nonseq_until_beta(engine, beta_threshold=0.02, batch=1000, max_samples=200_000, seed=9)
    assert out["beta"] < 0.02 and len(hist) == out["samples"] // 1000
    assert abs(out["plc"] - 0.0846) < 0.01 and 12 < out["edns"] < 17
