"""GPU parity of the analytical kernels (COPT, indices, F&D, Markov) -- 1e-9 relative or tighter."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-9   # north_star: COPT-derived LOLE/EENS within 1e-9 relative of the analytical result


def _for(mttf, mttr):
    lam = 1.0 / mttf; mu = 1.0 / mttr
    return lam / (lam + mu)


@pytest.mark.parametrize("step", [1.0, 10.0, 7.0, 0.5])
def test_copt_table_bit_exact(engine, rts, step):
    q = _for(rts["mttf"], rts["mttr"])
    g = engine.copt(rts["cap"], q, step)
    o = O.copt_build(rts["cap"], q, step)
    assert len(g) == len(o)
    assert np.array_equal(g, o)            # same FP64 operations in the same order


def test_run_analytical_known_answers(engine, rts):
    from powersystemsreliabilityassessment_b200 import Generator, LoadModel, run_analytical
    gens = [Generator(i + 1, c, a, b) for i, (c, a, b) in enumerate(zip(rts["cap"], rts["mttf"], rts["mttr"]))]
    for load, step, lole_ref, eue_ref in (
            (rts["load_mw"], 1.0, 9.3941103566, 1176.291677),
            (rts["load_int"].astype(float), 1.0, 9.3677375218, 1176.181257),
            (rts["load_mw"], 10.0, 9.4204746080, 1177.243237)):
        r = run_analytical(gens, LoadModel(load), step_size=step, engine=engine)
        lo, eo, _ = O.analytical(rts["cap"], _for(rts["mttf"], rts["mttr"]), load, step)
        assert r.method == "Analytical" and len(r.convergence_history) == 0
        assert abs(r.lole_hours_yr - lo) <= RTOL * lo and abs(r.eue_mwh_yr - eo) <= RTOL * eo
        assert abs(r.lole_hours_yr - lole_ref) < 5e-10 * 10 and abs(r.eue_mwh_yr - eue_ref) < 1e-6


def test_indices_overload_branch(engine):
    """Load above installed capacity hits the idx < 1 branch of PSA.jl:152-159."""
    cap = np.array([50.0, 50.0, 56.0, 100.0]); q = np.array([0.02, 0.02, 0.04, 0.05])
    load = np.array([300.0, 256.0, 255.9, 120.0, 10.0, 0.0, 270.5])
    p = engine.copt(cap, q, 10.0)
    lo, eo, po = O.analytical(cap, q, load, 10.0)
    assert np.array_equal(p, po)
    l, e = engine.copt_indices(p, 10.0, 256.0, load)
    assert abs(l - lo) <= RTOL * lo and abs(e - eo) <= RTOL * eo


def test_gaa_demo_strict_indices(engine):
    """generating_adequacy_assessment.jl:154-190 demo: LOLE 200.7899 h/yr, EUE 4930.1346 MWh/yr."""
    probs = O.gaa_build([50, 50, 56, 100], [0.02, 0.02, 0.04, 0.05], 10.0)
    ldc = np.array([200.0 - (100.0 / 8760) * (h - 1) for h in range(1, 8761)])
    l, e = engine.copt_indices_strict(probs, 10.0, ldc)
    lo, eo = O.gaa_indices(probs, 10.0, ldc)
    assert abs(l - lo) <= RTOL * lo and abs(e - eo) <= RTOL * eo
    assert abs(l - 200.789852160021) < 1e-7 and abs(e - 4930.134560000009) < 1e-6


def test_fd_recursion(engine, rts):
    from powersystemsreliabilityassessment_b200 import evaluate_risk
    P, F = engine.fd_recursion([16.0, 16.0], [4380.0, 4380.0], [89.39, 89.39])
    Po, Fo = O.fd_build([16.0, 16.0], [4380.0, 4380.0], [89.39, 89.39])
    assert np.array_equal(P, Po) and np.array_equal(F, Fo)
    lole, lolf, lold = evaluate_risk(P, F, 20.0, 32.0)
    assert abs(lole - 346.9045) < 1e-4 and abs(lolf - 3.8416) < 1e-4 and abs(lold - 90.3022) < 1e-4
    assert (lole, lolf, lold) == O.fd_evaluate(Po, Fo, 20.0, 32.0)
    P, F = engine.fd_recursion(rts["cap"], rts["mttf"] + rts["mttr"], rts["mttr"])
    Po, Fo = O.fd_build(rts["cap"], rts["mttf"] + rts["mttr"], rts["mttr"])
    assert len(P) == 3406 and np.array_equal(P, Po) and np.array_equal(F, Fo)


def test_markov2_and_dtmc(engine):
    m = engine.markov2(1000.0, 50.0, 1.0, 200)
    mo = O.markov2(1000.0, 50.0, 1.0, 200)
    assert np.allclose(m, mo, rtol=RTOL, atol=0)
    assert abs(m[-1] - 0.047333337093) < 1e-11
    rng = np.random.default_rng(42)
    mttf = [1000.0, 1200.0, 800.0, 1500.0, 2000.0]; mttr = [50.0, 60.0, 40.0, 20.0, 100.0]
    cap = [100.0, 100.0, 50.0, 200.0, 150.0]
    r = rng.random((1000, 5)); r[100:140] *= 0.01
    a = engine.dtmc_capacity(mttf, mttr, cap, r)
    ao = O.dtmc_capacity(mttf, mttr, cap, r)
    assert np.array_equal(a, ao) and a.min() < 600


def test_failure_time_experiment(engine):
    """Markov_process.jl:39-60: injected rand() and sampler streams, bit-exact vs the literal loop; exponential shape."""
    rng = np.random.default_rng(8)
    r = rng.random((64, 300)); r[:5] = 0.5                      # five components never fail within max_time = 200
    ft = engine.failure_times(0.01, 64, dt=1.0, max_time=200.0, uniforms=r)
    ref = O.failure_times(0.01, 1.0, 200.0, 64, uniforms=r)
    assert np.array_equal(ft, ref[ref >= 0]) and len(ft) <= 59
    with pytest.raises(Exception):
        engine.failure_times(0.01, 64, max_time=200.0, uniforms=r[:, :50] * 0 + 0.5)
    ft = engine.failure_times(1e-3, 10000, dt=1.0, max_time=5000.0, seed=42)     # the script's parameters
    ref = O.failure_times(1e-3, 1.0, 5000.0, 10000, seed=42)
    assert np.array_equal(ft, ref[ref >= 0])
    assert abs(len(ft) / 10000 - (1 - (1 - 1e-3) ** 5001)) < 0.01 and abs(ft.mean() - 966.0) < 30     # truncated geometric mean
