"""GPU parity of the sequential chronological MC (psra_seq_mc / psra_seq_eval_injected) against the
CPU oracle's literal restatement of PSA.jl:214-269 -- bit-exact per-year integers."""
import numpy as np
import pytest

from helpers import draws_needed, injected_durations
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _check_vs_literal(engine, cap, mttf, mttr, load_int, nchains, ypc, seed, dur_scale=1.0):
    rng = np.random.default_rng(seed)
    H = len(load_int)
    K = draws_needed(mttf * dur_scale, mttr * dur_scale, H * ypc)
    dur = injected_durations(rng, mttf, mttr, nchains, K, dur_scale)
    engine.set_system(cap, mttf, mttr)
    engine.set_load(load_int)
    r = engine.seq_eval_injected(dur, years_per_chain=ypc, fail_count=True)
    lol = []; ens = []; ent = []
    for c in range(nchains):
        a, b, e, used = O.seq_literal(cap, load_int.astype(np.float64), ypc, dur[c])
        assert used.max() <= K
        lol.append(a); ens.append(b); ent.append(e)
    lol = np.concatenate(lol); ens = np.concatenate(ens); ent = np.concatenate(ent)
    assert np.array_equal(r.lol_hours.astype(np.float64), lol)
    assert np.array_equal(r.raw["ens_fp_vector"].astype(np.float64), ens)
    assert np.array_equal(r.entries.astype(np.float64), ent)
    assert r.raw["sum_lol_hours"] == int(lol.sum())
    assert r.raw["sum_ens_fp"] == int(ens.sum())
    assert r.raw["sum_entries"] == int(ent.sum())
    assert r.raw["sum_lol_sq"] == int((lol.astype(np.int64) ** 2).sum())
    assert r.raw["sum_ens_sq"] == sum(int(x) ** 2 for x in ens)
    assert int(r.fail_count.sum()) == int(lol.sum())
    return lol


def test_injected_rts79_independent_years(engine, rts):
    lol = _check_vs_literal(engine, rts["cap"], rts["mttf"], rts["mttr"], rts["load_int"], 48, 1, 1)
    assert lol.sum() > 0          # the comparison is not vacuous


def test_injected_rts79_chain_carry(engine, rts):
    """State carries across the years of a chain exactly like PSA.jl:223-266."""
    _check_vs_literal(engine, rts["cap"], rts["mttf"], rts["mttr"], rts["load_int"], 6, 8, 2)


def test_injected_stressed_system(engine, rts):
    """Short MTTF/MTTR: many toggles per hour slot, same-hour multi-toggles, frequent loss."""
    _check_vs_literal(engine, rts["cap"], rts["mttf"], rts["mttr"], rts["load_int"], 8, 1, 3, dur_scale=0.02)


def test_injected_8760_hours_ragged_word(engine, rts):
    """H = 8760 (run_full_comparison.jl:19) is not a multiple of the 32-hour timeline word."""
    rng = np.random.default_rng(5)
    load = np.rint(1100 + 500 * np.sin((np.arange(1, 8761) - 2000) / 8760 * 2 * np.pi)
                   + 100 * rng.standard_normal(8760)).clip(0).astype(np.int32)
    cap = np.array([400, 400, 300, 300, 150, 150, 50, 50], dtype=np.float64)
    mttf = np.array([1100, 1100, 1200, 1200, 900, 900, 500, 500], dtype=np.float64)
    mttr = np.array([50, 50, 60, 60, 40, 40, 20, 20], dtype=np.float64)
    lol = _check_vs_literal(engine, cap, mttf, mttr, load, 16, 3, 7)
    assert lol.sum() > 0


def test_injected_many_units_generic_path(engine, rts):
    """U = 96 > 32 exercises the lane-strided unit loop and shared-memory unit states."""
    cap = np.tile(rts["cap"], 3); mttf = np.tile(rts["mttf"], 3); mttr = np.tile(rts["mttr"], 3)
    load = np.rint(3.05 * rts["load_mw"]).astype(np.int32)
    _check_vs_literal(engine, cap, mttf, mttr, load, 6, 1, 11)
    _check_vs_literal(engine, cap, mttf, mttr, load, 3, 2, 12)


def test_injected_tiny_and_edge_shapes(engine):
    cap = np.array([10.0, 5.0]); mttf = np.array([30.0, 20.0]); mttr = np.array([10.0, 15.0])
    for H in (1, 31, 32, 33, 100):
        load = np.full(H, 12, dtype=np.int32)
        _check_vs_literal(engine, cap, mttf, mttr, load, 4, 5, 100 + H)


def test_injected_overflow_is_reported(engine, rts):
    from powersystemsreliabilityassessment_b200 import PsraError
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"])
    engine.set_load(rts["load_int"])
    dur = np.full((1, 32, 4), 10.0)
    with pytest.raises(PsraError) as e:
        engine.seq_eval_injected(dur)
    assert e.value.code == -3


@pytest.mark.parametrize("init_mode", [0, 1])
def test_philox_mode_bit_exact_vs_oracle(engine, rts, init_mode):
    """Sampler mode: the oracle runs the literal hour loop with the same Philox/neg-log spec."""
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"])
    engine.set_load(rts["load_int"])
    r = engine.seq_mc(64, seed=1234, year0=128, init_mode=init_mode, per_year=True)
    lol, ens, ent = O.seq_philox(rts["cap"], rts["mttf"], rts["mttr"], rts["load_int"].astype(np.float64),
                                 1234, 128, 64, 1, init_mode)
    assert np.array_equal(r.lol_hours.astype(np.float64), lol)
    assert np.array_equal(r.raw["ens_fp_vector"].astype(np.float64), ens)
    assert np.array_equal(r.entries.astype(np.float64), ent)
    assert lol.sum() > 0


def test_philox_chain_mode_bit_exact(engine, rts):
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"])
    engine.set_load(rts["load_int"])
    r = engine.seq_mc(40, seed=7, year0=20, init_mode=0, years_per_chain=10, per_year=True)
    lol, ens, ent = O.seq_philox(rts["cap"], rts["mttf"], rts["mttr"], rts["load_int"].astype(np.float64),
                                 7, 2, 4, 10, 0)
    assert np.array_equal(r.lol_hours.astype(np.float64), lol)
    assert np.array_equal(r.raw["ens_fp_vector"].astype(np.float64), ens)
    assert np.array_equal(r.entries.astype(np.float64), ent)


def test_philox_generic_path_bit_exact(engine, rts):
    cap = np.tile(rts["cap"], 2); mttf = np.tile(rts["mttf"], 2); mttr = np.tile(rts["mttr"], 2)
    load = np.rint(2.05 * rts["load_mw"]).astype(np.int32)
    engine.set_system(cap, mttf, mttr); engine.set_load(load)
    r = engine.seq_mc(24, seed=99, init_mode=1, per_year=True)
    lol, ens, ent = O.seq_philox(cap, mttf, mttr, load.astype(np.float64), 99, 0, 24, 1, 1)
    assert np.array_equal(r.lol_hours.astype(np.float64), lol)
    assert np.array_equal(r.raw["ens_fp_vector"].astype(np.float64), ens)
    assert np.array_equal(r.entries.astype(np.float64), ent)


def test_sharding_invariance_and_segments(engine, rts):
    """Per-year integers do not depend on how years are split over calls (= GPUs) nor on the
    launch geometry (segment length, warps per block)."""
    from powersystemsreliabilityassessment_b200 import Engine
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(rts["load_int"])
    full = engine.seq_mc(4096, seed=5, per_year=True)
    a = engine.seq_mc(1024, seed=5, year0=0, per_year=True)
    b = engine.seq_mc(3072, seed=5, year0=1024, per_year=True)
    assert np.array_equal(full.lol_hours, np.concatenate([a.lol_hours, b.lol_hours]))
    assert np.array_equal(full.raw["ens_fp_vector"], np.concatenate([a.raw["ens_fp_vector"], b.raw["ens_fp_vector"]]))
    for k in ("sum_lol_hours", "sum_ens_fp", "sum_entries", "sum_lol_sq", "sum_ens_sq", "years_with_loss"):
        assert full.raw[k] == a.raw[k] + b.raw[k]
    for seg, wpb, gen, unp in ((8736, 4, False, False), (1120, 16, False, False), (320, 8, False, False), (32, 2, False, False),
                               (2208, 24, False, False), (8736, 4, True, False), (1120, 16, True, False), (320, 8, True, False),
                               (8736, 32, False, True), (8736, 7, False, True), (8736, 32, False, 1), (8736, 32, False, 2),
                               (8736, 32, False, 5), (8736, 20, True, 6)):
        sb = unp if (unp is not True and unp is not False) else 0       # static_blocks variants of the single-segment kernel
        with Engine(seg_hours=seg, warps_per_block=wpb, force_generic=gen and sb == 0, unpacked_words=unp is True,
                    static_blocks=sb) as e2:
            e2.set_system(rts["cap"], rts["mttf"], rts["mttr"]); e2.set_load(rts["load_int"])
            r2 = e2.seq_mc(4096, seed=5, per_year=True, group=10)
            assert np.array_equal(full.lol_hours, r2.lol_hours)
            assert np.array_equal(full.raw["ens_fp_vector"], r2.raw["ens_fp_vector"])
            assert np.array_equal(full.entries, r2.entries)
            assert np.array_equal(r2.group_lol[:409], full.lol_hours[:4090].reshape(-1, 10).sum(1))
            assert r2.raw["events"] == full.raw["events"]


def test_convergence_history_computed_on_device(engine, rts):
    """convergence_history = cum_lole / y every 10 years (PSA.jl:263-265); ragged tail years are dropped."""
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(rts["load_int"])
    for years in (10, 37, 4096, 25_013):
        r = engine.seq_mc(years, seed=11, per_year=True, history=10)
        k = years // 10
        want = np.cumsum(r.lol_hours[: 10 * k].astype(np.float64)).reshape(k, 10)[:, -1] / (10.0 * np.arange(1, k + 1))
        assert r.history.shape == (k,) and np.array_equal(r.history, want)
    assert engine.seq_mc(9, seed=11, history=10).history.shape == (0,)
    # long runs are cut into several launches whose history ranges are scanned / read back while the next one computes
    for years, ypc in ((2_500_003, 1), (1_200_000, 4)):
        r = engine.seq_mc(years, seed=3, per_year=True, history=10, group=10, years_per_chain=ypc) if years % ypc == 0 else \
            engine.seq_mc(years, seed=3, per_year=True, history=10, group=10)
        one = engine.seq_mc(years, seed=3, per_year=True, group=10, years_per_chain=ypc if years % ypc == 0 else 1)
        k = years // 10
        want = np.cumsum(r.lol_hours[: 10 * k].astype(np.float64)).reshape(k, 10)[:, -1] / (10.0 * np.arange(1, k + 1))
        assert np.array_equal(r.history, want) and np.array_equal(r.lol_hours, one.lol_hours)
        assert np.array_equal(r.group_lol, one.group_lol) and np.array_equal(r.raw["ens_fp_vector"], one.raw["ens_fp_vector"])
        for key in ("sum_lol_hours", "sum_ens_fp", "sum_entries", "sum_lol_sq", "sum_ens_sq", "years_with_loss", "events"):
            assert r.raw[key] == one.raw[key]
    import powersystemsreliabilityassessment_b200 as P
    gens = [P.Generator(i + 1, c, f, m) for i, (c, f, m) in enumerate(zip(rts["cap"], rts["mttf"], rts["mttr"]))]
    res = P.run_sequential_mc(gens, P.LoadModel(rts["load_int"].astype(np.float64)), 1000, seed=11, engine=engine)
    ref = engine.seq_mc(1000, seed=11, per_year=True)
    assert res.method == "Sequential MC" and len(res.convergence_history) == 100
    assert res.convergence_history[-1] == ref.lol_hours.sum() / 1000.0 == res.lole_hours_yr


def test_fast_and_generic_kernels_agree_in_chain_mode(engine, rts):
    """seq_fast.cu (wave / prefix-sum formulation) and seq_mc.cu (FP64 residual recurrence) are two
    independent implementations of the same chain; multi-year chains exercise the pending list."""
    from powersystemsreliabilityassessment_b200 import Engine
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(rts["load_int"])
    for init_mode, ypc in ((0, 5), (1, 25)):
        a = engine.seq_mc(3000, seed=17, init_mode=init_mode, years_per_chain=ypc, per_year=True, fail_count=True)
        with Engine(force_generic=True, seg_hours=4384) as g:
            g.set_system(rts["cap"], rts["mttf"], rts["mttr"]); g.set_load(rts["load_int"])
            b = g.seq_mc(3000, seed=17, init_mode=init_mode, years_per_chain=ypc, per_year=True, fail_count=True)
        assert np.array_equal(a.lol_hours, b.lol_hours) and np.array_equal(a.entries, b.entries)
        assert np.array_equal(a.raw["ens_fp_vector"], b.raw["ens_fp_vector"])
        assert np.array_equal(a.fail_count, b.fail_count)
        assert a.raw["events"] == b.raw["events"] and a.lol_hours.sum() > 0


def test_small_unit_counts_and_short_years_fast_path(engine):
    """U < 32 and tiny H through the sampler path vs the oracle's literal loop."""
    cap = np.array([10.0, 5.0, 7.0]); mttf = np.array([30.0, 20.0, 3.0]); mttr = np.array([10.0, 15.0, 2.0])
    for H in (1, 31, 33, 100, 1000):
        load = np.full(H, 14, dtype=np.int32)
        engine.set_system(cap, mttf, mttr); engine.set_load(load)
        for ypc in (1, 4):
            r = engine.seq_mc(32 * ypc, seed=H, init_mode=1, years_per_chain=ypc, per_year=True)
            lol, ens, ent = O.seq_philox(cap, mttf, mttr, load.astype(np.float64), H, 0, 32, ypc, 1)
            assert np.array_equal(r.lol_hours.astype(np.float64), lol)
            assert np.array_equal(r.raw["ens_fp_vector"].astype(np.float64), ens)
            assert np.array_equal(r.entries.astype(np.float64), ent)


def test_statistics_match_analytical_rts79(engine, rts):
    """Indices within the combined 95 % CI of the analytical COPT value (PSA.jl:293-294 benchmark):
    LOLE 9.3677 h/yr, EUE 1176.18 MWh/yr for the integer load curve."""
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(rts["load_int"])
    r = engine.seq_mc(400_000, seed=42)
    assert abs(r.lole - 9.3677375218) < 3.0 * r.lole_se + 1e-9
    assert abs(r.eens - 1176.181257) < 3.0 * r.eens_se + 1e-9
    assert 1.7 < r.lolf < 2.2 and 4.3 < r.lold < 5.4      # BASELINE.md ballparks: 1.90 occ/yr, 4.86 h
    assert 0.50 < r.p_loss_year < 0.60


def test_fixed_point_scale_int32_timeline(engine, rts):
    """fp_scale = 16: installed capacity 54 480 and peak load 45 600 exceed int16, so the sampler kernel
    takes its int32 timeline / int32 load-curve variant; ENS is then in 1/16 MWh."""
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"], fp_scale=16.0)
    load = engine.set_load(rts["load_mw"])                     # the grid values the library uses (ceil rule, api._fixed_load)
    assert np.abs(load - 16 * rts["load_mw"]).max() < 1.0 and (load >= 16 * rts["load_mw"] - 1e-9).all()
    r = engine.seq_mc(96, seed=321, init_mode=1, per_year=True)
    lol, ens, ent = O.seq_philox(16 * rts["cap"], rts["mttf"], rts["mttr"], load.astype(np.float64), 321, 0, 96, 1, 1)
    assert np.array_equal(r.lol_hours.astype(np.float64), lol)
    assert np.array_equal(r.raw["ens_fp_vector"].astype(np.float64), ens)
    assert np.array_equal(r.entries.astype(np.float64), ent)
    assert np.allclose(r.ens, ens / 16.0) and lol.sum() > 0
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"], fp_scale=1.0)


def test_matlab_discretisation_mode(engine, rts):
    """SURVEY a-8: round(TTF) / ceil(TTR), all UP each year (Montecarlo_seq/seq_mcsampling.m:40-74), both
    kernels vs the oracle's literal restatement of the MATLAB sampler evaluated at HL1."""
    from powersystemsreliabilityassessment_b200 import DISC_MATLAB, INIT_ALL_UP, Engine
    load = rts["load_int"]
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(load)
    r = engine.seq_mc(96, seed=77, year0=32, init_mode=INIT_ALL_UP | DISC_MATLAB, per_year=True)
    lol, ens, ent = O.seq_matlab_philox(rts["cap"], rts["mttf"], rts["mttr"], load.astype(np.float64), 77, 32, 96)
    assert np.array_equal(r.lol_hours.astype(np.float64), lol)
    assert np.array_equal(r.raw["ens_fp_vector"].astype(np.float64), ens)
    assert np.array_equal(r.entries.astype(np.float64), ent) and lol.sum() > 0
    with Engine(force_generic=True) as g:
        g.set_system(rts["cap"], rts["mttf"], rts["mttr"]); g.set_load(load)
        r2 = g.seq_mc(96, seed=77, year0=32, init_mode=INIT_ALL_UP | DISC_MATLAB, per_year=True)
    assert np.array_equal(r.lol_hours, r2.lol_hours) and np.array_equal(r.raw["ens_fp_vector"], r2.raw["ens_fp_vector"])
    # short cycles: zero-hour up times and same-hour double toggles
    cap = np.array([10.0, 5.0, 7.0]); mttf = np.array([3.0, 0.4, 30.0]); mttr = np.array([1.5, 2.0, 0.2])
    ld = np.full(500, 14, dtype=np.int32)
    engine.set_system(cap, mttf, mttr); engine.set_load(ld)
    r3 = engine.seq_mc(64, seed=5, init_mode=INIT_ALL_UP | DISC_MATLAB, per_year=True)
    lol, ens, ent = O.seq_matlab_philox(cap, mttf, mttr, ld.astype(np.float64), 5, 0, 64)
    assert np.array_equal(r3.lol_hours.astype(np.float64), lol) and np.array_equal(r3.entries.astype(np.float64), ent)
    assert np.array_equal(r3.raw["ens_fp_vector"].astype(np.float64), ens)


def test_team_kernel_large_systems(engine, rts):
    """seq_team.cu (block per chain, > 32 units): bit-exact vs the oracle's literal loop and vs the generic
    kernel, incl. a unit count that is not a multiple of 32, multi-year chains and BASELINE config 5 (1024 units)."""
    from powersystemsreliabilityassessment_b200 import DISC_MATLAB, INIT_ALL_UP, Engine, rts79
    cap = np.tile(rts["cap"], 3)[:70]; mttf = np.tile(rts["mttf"], 3)[:70]; mttr = np.tile(rts["mttr"], 3)[:70]
    load = np.rint(2.2 * rts["load_mw"]).astype(np.int32)
    engine.set_system(cap, mttf, mttr); engine.set_load(load)
    for ypc, init in ((1, 1), (4, 0)):
        r = engine.seq_mc(24 * ypc, seed=31, init_mode=init, years_per_chain=ypc, per_year=True, fail_count=True)
        lol, ens, ent = O.seq_philox(cap, mttf, mttr, load.astype(np.float64), 31, 0, 24, ypc, init)
        assert np.array_equal(r.lol_hours.astype(np.float64), lol) and np.array_equal(r.entries.astype(np.float64), ent)
        assert np.array_equal(r.raw["ens_fp_vector"].astype(np.float64), ens) and lol.sum() > 0
        with Engine(force_generic=True) as g:
            g.set_system(cap, mttf, mttr); g.set_load(load)
            r2 = g.seq_mc(24 * ypc, seed=31, init_mode=init, years_per_chain=ypc, per_year=True, fail_count=True)
        assert np.array_equal(r.lol_hours, r2.lol_hours) and np.array_equal(r.fail_count, r2.fail_count)
        assert r.raw["events"] == r2.raw["events"]
    r = engine.seq_mc(16, seed=8, init_mode=INIT_ALL_UP | DISC_MATLAB, per_year=True)
    lol, ens, ent = O.seq_matlab_philox(cap, mttf, mttr, load.astype(np.float64), 8, 0, 16)
    assert np.array_equal(r.lol_hours.astype(np.float64), lol) and np.array_equal(r.raw["ens_fp_vector"].astype(np.float64), ens)
    # config 5: 1024 units, load x 37
    cap, mttf, mttr, load = rts79.synthetic_system(32, 37.0)
    engine.set_system(cap, mttf, mttr); engine.set_load(load)
    r = engine.seq_mc(12, seed=2024, year0=100, per_year=True)
    lol, ens, ent = O.seq_philox(cap, mttf, mttr, load.astype(np.float64), 2024, 100, 12, 1, 1)
    assert np.array_equal(r.lol_hours.astype(np.float64), lol) and np.array_equal(r.entries.astype(np.float64), ent)
    assert np.array_equal(r.raw["ens_fp_vector"].astype(np.float64), ens)
    big = engine.seq_mc(20000, seed=1)
    assert abs(big.lole - 8.033131) < 4 * big.lole_se          # analytical COPT value
    # seq_wide.cu (lane-level unit queue, the default for whole-year timelines) against seq_team.cu (wave scheduler)
    with Engine(force_team=True) as tm:
        tm.set_system(cap, mttf, mttr); tm.set_load(load)
        for disc in (0, DISC_MATLAB):
            w = engine.seq_mc(300, seed=77, init_mode=1 | disc, per_year=True, fail_count=True, group=10)
            t = tm.seq_mc(300, seed=77, init_mode=1 | disc, per_year=True, fail_count=True, group=10)
            assert np.array_equal(w.lol_hours, t.lol_hours) and np.array_equal(w.entries, t.entries)
            assert np.array_equal(w.raw["ens_fp_vector"], t.raw["ens_fp_vector"]) and np.array_equal(w.fail_count, t.fail_count)
            assert np.array_equal(w.group_lol, t.group_lol) and w.raw["events"] == t.raw["events"] and w.lol_hours.sum() > 0
            for k in ("sum_lol_hours", "sum_ens_fp", "sum_entries", "sum_lol_sq", "sum_ens_sq", "years_with_loss"):
                assert w.raw[k] == t.raw[k]


def test_high_transition_rate_short_segments(engine):
    """Units that toggle every few hours force the sampler kernel into many short ring segments (event lists
    are bounded); results stay bit-identical to the literal loop."""
    cap = np.array([40.0, 30.0, 30.0, 20.0, 20.0, 10.0, 10.0, 5.0])
    mttf = np.array([2.0, 3.0, 2.5, 4.0, 1.5, 6.0, 2.0, 1.0]); mttr = np.array([1.0, 0.5, 2.0, 1.0, 0.7, 3.0, 0.2, 1.0])
    rng = np.random.default_rng(3)
    load = rng.integers(60, 130, 8736).astype(np.int32)
    engine.set_system(cap, mttf, mttr); engine.set_load(load)
    for ypc, init in ((1, 1), (3, 0)):
        r = engine.seq_mc(12 * ypc, seed=55, init_mode=init, years_per_chain=ypc, per_year=True)
        lol, ens, ent = O.seq_philox(cap, mttf, mttr, load.astype(np.float64), 55, 0, 12, ypc, init)
        assert np.array_equal(r.lol_hours.astype(np.float64), lol) and np.array_equal(r.entries.astype(np.float64), ent)
        assert np.array_equal(r.raw["ens_fp_vector"].astype(np.float64), ens)
        assert r.raw["events"] > 30000 * 12 * ypc and lol.sum() > 100


@pytest.mark.gpu
def test_sampler_logarithm_draw_by_draw(engine):
    """The device logarithm / tick duration of single draws against the CPU specification: edge draws (0, 1, all ones,
    powers of two and their neighbours, the sqrt(1/2) reduction boundary), a stride over the whole 32-bit range and
    random draws; several means incl. durations below one tick and above 2^32 ticks."""
    rng = np.random.default_rng(5)
    edge = [0, 1, 2, 3, 0xFFFFFFFF, 0xFFFFFFFE, 0x80000000, 0x7FFFFFFF, 0x80000001, 0xB504F333, 0xB504F334, 0xB504F332,
            0x5A827999, 0x5A82799A, 0x00B504F3, 0x00000100, 0x000000FF, 0x01000000, 0x00FFFFFF]
    edge += [1 << k for k in range(32)] + [(1 << k) - 1 for k in range(1, 32)] + [(1 << k) + 1 for k in range(1, 32)]
    draws = np.concatenate([np.array(edge, dtype=np.uint64), np.arange(0, 1 << 32, 1021, dtype=np.uint64),
                            rng.integers(0, 1 << 32, 1 << 20, dtype=np.uint64)]).astype(np.uint32)
    for mean in (2940.0, 450.0, 20.0, 0.1, 1e-7, 3.0e5):
        t_gpu, e_gpu = engine.sampler_durations(mean, draws)
        t_cpu, e_cpu = O.sampler_durations(mean, draws)
        assert np.array_equal(e_gpu, e_cpu)
        assert np.array_equal(t_gpu, t_cpu)
    assert t_cpu.max() > (1 << 32) and O.sampler_durations(1e-7, draws)[0].min() == 1


@pytest.mark.gpu
def test_randomised_systems_sampler_path_vs_oracle(engine):
    """Seeded sweep over random systems through the sampler kernels (single-segment packed / unpacked, two-segment
    ring, multi-year chains, MATLAB discretisation): unit counts 1..32, ragged hour counts, short and long cycles,
    loads around the installed capacity -- every year bit-exact against the oracle's literal hour/unit loop."""
    from powersystemsreliabilityassessment_b200 import DISC_MATLAB
    rng = np.random.default_rng(20261017)
    checked = 0
    for case in range(36):
        U = int(rng.integers(1, 33))
        H = int(rng.choice([rng.integers(1, 200), rng.integers(200, 3000), rng.integers(3000, 9000), 8736, 8760]))
        scale = float(rng.choice([1.0, 1.0, 37.0, 400.0]))                    # large capacities leave the packed-word range
        cap = np.rint(rng.integers(1, 400, U) * scale).astype(np.float64)
        cyc = float(rng.choice([15.0, 60.0, 400.0, 2000.0]))
        mttf = rng.uniform(0.5, 1.5, U) * cyc
        mttr = rng.uniform(0.05, 0.6, U) * cyc
        lvl = rng.uniform(0.55, 1.0)
        load = np.rint(cap.sum() * lvl * (0.75 + 0.25 * np.sin(np.arange(H) / 24.0 * 2 * np.pi) * rng.uniform(0, 1))).astype(np.int32)
        ypc = int(rng.choice([1, 1, 1, 3]))
        init = int(rng.choice([0, 1]))
        disc = bool(rng.integers(0, 4) == 0)
        if disc:                                      # the MATLAB sampler (seq_mcsampling.m): independent years, all UP at hour 0
            ypc, init = 1, 0
        years = 8 * ypc
        try:
            engine.set_system(cap, mttf, mttr); engine.set_load(load)
            r = engine.seq_mc(years, seed=1000 + case, year0=ypc * 5, init_mode=init | (DISC_MATLAB if disc else 0),
                              years_per_chain=ypc, per_year=True)
        except Exception as ex:                       # the only legitimate refusal: transition rate too high for the list
            assert "transition rate too high" in str(ex) or "overflow" in str(ex).lower(), (case, ex)
            continue
        if disc:
            lol, ens, ent = O.seq_matlab_philox(cap, mttf, mttr, load.astype(np.float64), 1000 + case, 5, 8)
        else:
            lol, ens, ent = O.seq_philox(cap, mttf, mttr, load.astype(np.float64), 1000 + case, 5, 8, ypc, init)
        assert np.array_equal(r.lol_hours.astype(np.float64), lol), (case, U, H, ypc, init, disc)
        assert np.array_equal(r.raw["ens_fp_vector"].astype(np.float64), ens), (case, U, H, ypc, init, disc)
        assert np.array_equal(r.entries.astype(np.float64), ent), (case, U, H, ypc, init, disc)
        checked += 1
    assert checked >= 30


@pytest.mark.gpu
def test_unit_importance_weak_points(engine, rts):
    """Montecarlo_seq/seqMain.m:140-150,225-231 at HL1: hours with loss of load in which each unit is DOWN, against the
    oracle's literal hour/unit loop -- RTS-79 (both start modes, a multi-year chain, the MATLAB-free default
    discretisation), a ragged short year with frequent losses, and a 64-unit system (generic strided path)."""
    load = rts["load_int"]
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(load)
    for init, ypc, years, y0 in ((1, 1, 96, 32), (0, 1, 64, 0), (1, 4, 64, 8)):
        imp, cnt, r = engine.seq_unit_importance(years, seed=31, year0=y0, init_mode=init, years_per_chain=ypc, per_year=True)
        lol, ens, ent, ref = O.seq_philox_importance(rts["cap"], rts["mttf"], rts["mttr"], load.astype(np.float64), 31,
                                                     y0 // ypc, years // ypc, ypc, init)
        assert np.array_equal(r.lol_hours.astype(np.float64), lol) and np.array_equal(r.raw["ens_fp_vector"].astype(np.float64), ens)
        assert np.array_equal(cnt.astype(np.float64), ref) and ref.sum() > 0
        assert np.allclose(imp, ref / lol.sum())
        assert imp[21] > 0.5 and imp[22] > 0.5            # the two 400 MW units dominate the loss hours
    # same per-year integers as the production kernel
    fast = engine.seq_mc(96, seed=31, year0=32, per_year=True)
    _, _, r = engine.seq_unit_importance(96, seed=31, year0=32, per_year=True)
    assert np.array_equal(fast.lol_hours, r.lol_hours) and np.array_equal(fast.raw["ens_fp_vector"], r.raw["ens_fp_vector"])
    # short ragged year, frequent losses, same-hour double toggles
    cap = np.array([10.0, 5.0, 7.0]); mttf = np.array([30.0, 20.0, 3.0]); mttr = np.array([10.0, 15.0, 2.0])
    ld = np.full(1000, 14, dtype=np.int32)
    engine.set_system(cap, mttf, mttr); engine.set_load(ld)
    imp, cnt, r = engine.seq_unit_importance(64, seed=3, per_year=True)
    lol, ens, ent, ref = O.seq_philox_importance(cap, mttf, mttr, ld.astype(np.float64), 3, 0, 64, 1, 1)
    assert np.array_equal(cnt.astype(np.float64), ref) and np.array_equal(r.lol_hours.astype(np.float64), lol)
    # 64 units: lane-strided generic path
    cap2 = np.tile(rts["cap"], 2); mttf2 = np.tile(rts["mttf"], 2); mttr2 = np.tile(rts["mttr"], 2)
    ld2 = np.rint(2.05 * rts["load_mw"]).astype(np.int32)
    engine.set_system(cap2, mttf2, mttr2); engine.set_load(ld2)
    imp, cnt, r = engine.seq_unit_importance(48, seed=17, per_year=True)
    lol, ens, ent, ref = O.seq_philox_importance(cap2, mttf2, mttr2, ld2.astype(np.float64), 17, 0, 48, 1, 1)
    assert np.array_equal(r.lol_hours.astype(np.float64), lol) and np.array_equal(cnt.astype(np.float64), ref) and ref.sum() > 0
    with pytest.raises(Exception):
        engine.seq_unit_importance(48, seed=17, years_per_chain=4)
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(load)
