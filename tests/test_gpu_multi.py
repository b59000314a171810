"""Multi-GPU inside libpsra_b200.so (psra_config.ngpus, csrc/multi.cu): ONE host call shards the years / samples over
the devices of the process and combines the integers with ncclAllReduce.  The per-year integers are keyed on the
global year, so 1 / 2 / 4 / 8 devices must give identical results (the call site is
run_sequential_mc(gens, load, years), GeneratingAdequacy/run_full_comparison.jl:32).  Needs >= 2 visible devices
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`); skipped on a one-GPU box."""
import numpy as np
import pytest

import powersystemsreliabilityassessment_b200 as P
from powersystemsreliabilityassessment_b200 import Engine, rts79

pytestmark = pytest.mark.gpu

SUMS = ("years", "sum_lol_hours", "sum_ens_fp", "sum_entries", "sum_lol_sq", "sum_ens_sq", "years_with_loss", "events")


def _ndev():
    import torch
    return torch.cuda.device_count()


def _counts():
    return [g for g in (2, 3, 4, 8) if g <= _ndev()]


@pytest.fixture(scope="module")
def need2():
    if _ndev() < 2:
        pytest.skip("needs >= 2 CUDA devices")


def test_ngpus_beyond_the_visible_devices_is_refused():
    with pytest.raises(P.PsraError) as e:
        Engine(ngpus=_ndev() + 1)
    assert e.value.code == -1 and "device" in str(e.value)


def test_sequential_identical_integers_for_any_device_count(need2, rts):
    """RTS-79 (seq_fast.cu): sums, per-year vectors, per-hour failure counts, group sums, the running-mean history and
    the VaR / CVaR from the all-reduced ENS histogram are those of the one-device run; year count not divisible by
    anything, year0 != 0."""
    n, y0 = 1_000_003, 70
    with Engine() as one:
        one.set_system(rts["cap"], rts["mttf"], rts["mttr"]); one.set_load(rts["load_int"])
        a = one.seq_mc(n, seed=13, year0=y0, per_year=True, fail_count=True, group=10, history=10, tail_hist=True)
        ta = one.tail(None, alphas=(0.5, 0.95, 0.99))
    for G in _counts():
        with Engine(ngpus=G) as eng:
            eng.set_system(rts["cap"], rts["mttf"], rts["mttr"]); eng.set_load(rts["load_int"])
            b = eng.seq_mc(n, seed=13, year0=y0, per_year=True, fail_count=True, group=10, history=10, tail_hist=True)
            for k in SUMS:
                assert a.raw[k] == b.raw[k], (G, k)
            assert np.array_equal(a.lol_hours, b.lol_hours) and np.array_equal(a.raw["ens_fp_vector"], b.raw["ens_fp_vector"])
            assert np.array_equal(a.entries, b.entries) and np.array_equal(a.fail_count, b.fail_count)
            assert np.array_equal(a.group_lol, b.group_lol) and np.array_equal(a.history, b.history)
            assert eng.tail(None, alphas=(0.5, 0.95, 0.99)) == ta
            # accumulators only (the bench's call), and fewer years than devices
            c = eng.seq_mc(n, seed=13, year0=y0)
            assert all(a.raw[k] == c.raw[k] for k in SUMS)
            d = eng.seq_mc(1, seed=13, year0=y0, per_year=True, tail_hist=True)
            assert d.lol_hours[0] == a.lol_hours[0] and d.raw["years"] == 1


def test_config5_and_chains_shard_over_devices(need2):
    """1024 units (seq_wide.cu) and three-year chains of a 32-unit system (ring kernel): shards are whole chains and
    whole history groups."""
    cap, mttf, mttr, load = rts79.synthetic_system(32, 37.0)
    with Engine() as one:
        one.set_system(cap, mttf, mttr); one.set_load(load)
        a = one.seq_mc(3001, seed=2, per_year=True, group=10, history=10, fail_count=True)
    c1, f1, r1 = rts79.units()
    l1 = rts79.load_curve_int()
    with Engine() as one:
        one.set_system(c1, f1, r1); one.set_load(l1)
        ch = one.seq_mc(3 * 2001, seed=4, init_mode=0, years_per_chain=3, per_year=True, group=10, history=10)
    for G in _counts():
        with Engine(ngpus=G) as eng:
            eng.set_system(cap, mttf, mttr); eng.set_load(load)
            b = eng.seq_mc(3001, seed=2, per_year=True, group=10, history=10, fail_count=True)
            assert all(a.raw[k] == b.raw[k] for k in SUMS)
            assert np.array_equal(a.lol_hours, b.lol_hours) and np.array_equal(a.history, b.history) and np.array_equal(a.fail_count, b.fail_count)
            eng.set_system(c1, f1, r1); eng.set_load(l1)
            cb = eng.seq_mc(3 * 2001, seed=4, init_mode=0, years_per_chain=3, per_year=True, group=10, history=10)
            assert all(ch.raw[k] == cb.raw[k] for k in SUMS)
            assert np.array_equal(ch.lol_hours, cb.lol_hours) and np.array_equal(ch.history, cb.history) and np.array_equal(ch.group_lol, cb.group_lol)


def test_non_sequential_shards_over_devices(need2, rts):
    with Engine() as one:
        one.set_system(rts["cap"], rts["mttf"], rts["mttr"]); one.set_load(rts["load_int"])
        a = one.nonseq_mc(1_000_037, seed=3, sample0=11, per_sample=True, states=True, group=100, history=100)
    for G in _counts():
        with Engine(ngpus=G) as eng:
            eng.set_system(rts["cap"], rts["mttf"], rts["mttr"]); eng.set_load(rts["load_int"])
            b = eng.nonseq_mc(1_000_037, seed=3, sample0=11, per_sample=True, states=True, group=100, history=100)
            assert a["raw"] == b["raw"]
            for k in ("lol_hours", "ens", "cap", "states", "group_lol", "history"):
                assert np.array_equal(a[k], b[k]), (G, k)


def test_drop_in_call_uses_all_devices(need2, rts):
    """The reference's call, run_sequential_mc(gens, load, years), with an engine that spans every visible device."""
    gens = [P.Generator(i + 1, float(c), float(a), float(b)) for i, (c, a, b) in enumerate(zip(rts["cap"], rts["mttf"], rts["mttr"]))]
    lm = P.LoadModel(rts["load_int"].astype(np.float64))
    with Engine() as one, Engine(ngpus=_ndev()) as many:
        r1, i1 = P.run_sequential_mc(gens, lm, 2_000_000, seed=9, engine=one, details=True)
        r2, i2 = P.run_sequential_mc(gens, lm, 2_000_000, seed=9, engine=many, details=True)
        assert r1.lole_hours_yr == r2.lole_hours_yr and r1.eue_mwh_yr == r2.eue_mwh_yr
        assert np.array_equal(r1.convergence_history, r2.convergence_history) and i1.raw == i2.raw
        n1 = P.run_non_sequential_mc(gens, lm, 3_000_000, seed=9, engine=one)
        n2 = P.run_non_sequential_mc(gens, lm, 3_000_000, seed=9, engine=many)
        assert n1.lole_hours_yr == n2.lole_hours_yr and np.array_equal(n1.convergence_history, n2.convergence_history)
