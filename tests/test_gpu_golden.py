"""The CUDA path against the committed golden fixtures (tests/golden/*, written by scripts/make_golden.py from the
pinned oracle): the fixtures travel to the GPU box, where /root/reference does not exist."""
import os

import numpy as np
import pytest

from helpers import injected_durations

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_copt_table_golden(engine, rts):
    lam = 1.0 / rts["mttf"]; mu = 1.0 / rts["mttr"]; q = lam / (lam + mu)
    probs = engine.copt(rts["cap"], q, 10.0)
    assert np.array_equal(probs, np.load(os.path.join(G, "rts79_copt_step10.npy")))       # bit-identical FP64 table


def test_injected_literal_golden(engine, rts):
    g = np.load(os.path.join(G, "seq_literal_seed123.npz"))
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(rts["load_int"])
    r = engine.seq_eval_injected(g["dur"][None, :, :], years_per_chain=3)
    assert np.array_equal(r.lol_hours.astype(np.float64), g["lol"])
    assert np.array_equal(r.raw["ens_fp_vector"].astype(np.float64), g["eue"])
    assert np.array_equal(r.entries.astype(np.float64), g["ent"])
    # the fixture's input is what the committed generator script draws
    dur = injected_durations(np.random.default_rng(123), rts["mttf"], rts["mttr"], 1, 200)[0]
    assert np.array_equal(dur, g["dur"])


def test_sampler_years_golden(engine, rts):
    g = np.load(os.path.join(G, "seq_philox_seed42.npz"))
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(rts["load_int"])
    r = engine.seq_mc(64, seed=42, per_year=True)
    assert np.array_equal(r.lol_hours.astype(np.float64), g["lol"])
    assert np.array_equal(r.raw["ens_fp_vector"].astype(np.float64), g["eue"])
    assert np.array_equal(r.entries.astype(np.float64), g["ent"])


def test_sampler_draws_golden(engine):
    g = np.load(os.path.join(G, "sampler_draws.npz"))
    t, e = engine.sampler_durations(450.0, g["draws"])
    assert np.array_equal(e, g["e_bits"]) and np.array_equal(t, g["ticks_450"])
    t, _ = engine.sampler_durations(2940.0, g["draws"])
    assert np.array_equal(t, g["ticks_2940"])


def test_nonsequential_samples_golden(engine, rts):
    g = np.load(os.path.join(G, "nonseq_philox_seed7.npz"))
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(rts["load_int"])
    r = engine.nonseq_mc(256, seed=7, sample0=1000, per_sample=True, states=True)
    assert np.array_equal(r["lol_hours"].astype(np.float64), g["lol"])
    assert np.array_equal(r["ens"].astype(np.float64), g["eue"])
    assert np.array_equal(r["states"].reshape(256, -1), g["states"])
