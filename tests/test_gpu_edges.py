"""Edge cases of the sequential / non-sequential path through the C ABI, against the CPU oracle: empty runs, systems that
never / always lose load, one unit, the unit counts where the kernel family changes (32 | 33, 2048 | 2049), zero-capacity
units, transition rates from one per hour to one per 10^7 hours, and the argument errors of the boundary."""
import numpy as np
import pytest

import powersystemsreliabilityassessment_b200 as P
from oracle import oracle as O
from powersystemsreliabilityassessment_b200 import Engine, rts79

pytestmark = pytest.mark.gpu


def _check(eng, cap, mttf, mttr, load, n, seed=3, year0=0, init=1):
    eng.set_system(cap, mttf, mttr); eng.set_load(load)
    r = eng.seq_mc(n, seed=seed, year0=year0, init_mode=init, per_year=True, fail_count=True)
    lol, ens, ent = O.seq_philox(cap, mttf, mttr, np.asarray(load, dtype=np.float64), seed, year0, n, 1, init)
    assert np.array_equal(r.lol_hours.astype(np.float64), lol), "LOL hours"
    assert np.array_equal(r.raw["ens_fp_vector"].astype(np.float64), ens), "ENS"
    assert np.array_equal(r.entries.astype(np.float64), ent), "entries"
    assert r.raw["sum_lol_hours"] == int(lol.sum()) and int(r.fail_count.sum()) == int(lol.sum())
    return r


def test_empty_run_and_no_loss_and_permanent_loss(engine, rts):
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(rts["load_int"])
    z = engine.seq_mc(0, seed=1)
    assert z.raw["years"] == 0 and z.raw["sum_lol_hours"] == 0 and z.raw["events"] == 0
    g = engine.nonseq_mc(0, seed=1)
    assert g["raw"]["samples"] == 0
    H = 8736
    # load 0: never a loss, no entries; load above the installed capacity: every hour lost, ONE entry per year (calnlc.m:30-32)
    r = _check(engine, rts["cap"], rts["mttf"], rts["mttr"], np.zeros(H, dtype=np.int32), 40)
    assert r.raw["sum_lol_hours"] == 0 and r.raw["years_with_loss"] == 0 and r.raw["events"] > 0
    r = _check(engine, rts["cap"], rts["mttf"], rts["mttr"], np.full(H, int(rts["cap"].sum()) + 1, dtype=np.int32), 40)
    assert (r.lol_hours == H).all() and (r.entries == 1).all() and r.lolf == 1.0
    # load equal to the installed capacity: the comparison is strict (PSA.jl:253) -- a loss only while a unit is down
    r = _check(engine, rts["cap"], rts["mttf"], rts["mttr"], np.full(H, int(rts["cap"].sum()), dtype=np.int32), 24)
    assert 0 < r.raw["sum_lol_hours"] < 24 * H


@pytest.mark.parametrize("U", [1, 2, 31, 32, 33, 64, 65, 2048, 2049])
def test_unit_counts_across_the_kernel_families(engine, U):
    """1 .. 32 units: seq_fast.cu; 33 .. 2048: seq_wide.cu (static / queue phases, padding lanes); 2049: the generic kernel."""
    rng = np.random.default_rng(U)
    cap = rng.integers(0, 60, U).astype(np.float64)          # zero-capacity units included
    cap[0] = 50.0
    mttf = rng.uniform(200.0, 3000.0, U); mttr = rng.uniform(10.0, 150.0, U)
    H = int(rng.choice([168, 1000, 8736]))
    load = np.rint(rng.uniform(0.75, 0.97, H) * cap.sum()).astype(np.int32)
    r = _check(engine, cap, mttf, mttr, load, 6 if U > 1000 else 24, seed=U, year0=5)
    assert r.lol_hours.sum() > 0


def test_extreme_transition_rates(engine):
    """Units that toggle about once per hour (many events per 32-hour word: the event lists of seq_fast.cu overflow and the
    library replays / falls back by itself) next to units that practically never fail (MTTF 10^7 h)."""
    rng = np.random.default_rng(4)
    for U in (8, 40):
        cap = rng.integers(5, 40, U).astype(np.float64)
        mttf = np.where(np.arange(U) % 2 == 0, rng.uniform(1.0, 3.0, U), 1.0e7)
        mttr = np.where(np.arange(U) % 2 == 0, rng.uniform(0.5, 2.0, U), 1.0e7)
        load = np.rint(rng.uniform(0.5, 0.8, 700) * cap.sum()).astype(np.int32)
        r = _check(engine, cap, mttf, mttr, load, 16, seed=9)
        assert r.raw["events"] > 16 * 700 * (U // 2) // 4 and r.lol_hours.sum() > 0
        r = _check(engine, cap, mttf, mttr, load, 8, seed=9, init=0)       # all-up start (the reference's)
    # one unit, one hour
    _check(engine, np.array([10.0]), np.array([2.0]), np.array([2.0]), np.array([5], dtype=np.int32), 64)


def test_argument_errors_of_the_boundary(rts):
    with Engine() as e:
        with pytest.raises(P.PsraError) as ei:
            e.seq_mc(10, seed=1)                                            # no system yet
        assert ei.value.code == -1
        e.set_system(rts["cap"], rts["mttf"], rts["mttr"])
        with pytest.raises(P.PsraError):
            e.seq_mc(10, seed=1)                                            # no load yet
        e.set_load(rts["load_int"])
        with pytest.raises(P.PsraError):
            e.seq_mc(10, seed=1, years_per_chain=3)                         # years not a multiple of the chain length
        with pytest.raises(P.PsraError):
            e.seq_mc(-1, seed=1)
        for bad in (dict(mttf=-1.0), dict(mttr=0.0), dict(mttf=float("nan")), dict(mttr=2.0e8)):
            f, r = rts["mttf"].copy(), rts["mttr"].copy()
            if "mttf" in bad: f[3] = bad["mttf"]
            else: r[3] = bad["mttr"]
            with pytest.raises(P.PsraError):
                e.set_system(rts["cap"], f, r)
        with pytest.raises((P.PsraError, ValueError)):
            e.set_system(-rts["cap"], rts["mttf"], rts["mttr"])
        with pytest.raises(ValueError):
            e.set_system(rts["cap"] + 0.5, rts["mttf"], rts["mttr"])        # not representable at fp_scale 1 (strict)
        # the engine is still usable after the refused calls
        e.set_system(rts["cap"], rts["mttf"], rts["mttr"]); e.set_load(rts["load_int"])
        assert e.seq_mc(100, seed=1).raw["years"] == 100
