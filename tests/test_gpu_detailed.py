"""GPU parity of the hourly-resampled MC with maintenance / LFU / energy-limited units (tail_risk.jl:12-91,
MCvsMarkovProcess.jl:210-284) against the oracle's literal restatement."""
import math

import numpy as np
import pytest

from oracle import oracle as O
import powersystemsreliabilityassessment_b200 as P

pytestmark = pytest.mark.gpu


def _system():
    """The 6-unit system of tail_risk.jl:148-158 with a seeded stand-in for its unseeded randn() load."""
    gens = [P.DetailedGenerator("Nuclear", 400.0, 0.02, 4), P.DetailedGenerator("Coal_A", 300.0, 0.04, 3),
            P.DetailedGenerator("Coal_B", 300.0, 0.04, 3), P.DetailedGenerator("Gas", 150.0, 0.05, 2),
            P.DetailedGenerator("Hydro_ELU", 200.0, 0.01, 2, 200.0 * 50.0), P.DetailedGenerator("Old_56", 56.0, 0.10, 0)]
    rng = np.random.default_rng(7)
    h = np.arange(1, 8761)
    base = np.maximum(0.0, 750.0 + 300.0 * np.sin((h - 2000) / 8760 * 2 * math.pi) + 50.0 * rng.standard_normal(8760))
    peaks = [base[(w - 1) * 168:min(w * 168, 8760)].max() for w in range(1, 53)]
    P.schedule_maintenance(gens, peaks)
    return gens, base, peaks


def _arrays(gens):
    return ([g.capacity for g in gens], [g.for_rate for g in gens], [g.scheduled_outage_start for g in gens],
            [g.maintenance_weeks for g in gens], [g.energy_limit for g in gens])


def test_schedule_maintenance_matches_oracle():
    gens, base, peaks = _system()
    ref = O.schedule_maintenance([g.capacity for g in gens], [g.maintenance_weeks for g in gens], peaks)
    assert [g.scheduled_outage_start for g in gens] == ref
    assert all(1 <= g.scheduled_outage_start <= 52 for g in gens if g.maintenance_weeks) and gens[5].scheduled_outage_start == 0


def test_injected_bit_exact(engine):
    gens, base, _ = _system()
    rng = np.random.default_rng(1)
    n = 6
    unif = rng.random((n, 8760, 6)); norm = rng.standard_normal((n, 8760))
    unif[:, 3000:3600, :4] *= 0.25            # outage-rich stretch: deficits and ELU exhaustion do occur
    lfu = base.max() * 0.05
    yl, hf = engine.detailed_eval_injected(gens, base, lfu, unif, norm)
    ryl, rhf = O.detailed_mc_injected(*_arrays(gens), base, lfu, unif, norm)
    assert np.array_equal(yl.astype(float), ryl) and np.array_equal(hf.astype(float), rhf)
    assert ryl.sum() > 50


def test_sampler_bit_exact_and_sharding(engine):
    gens, base, _ = _system()
    lfu = base.max() * 0.05
    yl, hf, _ = engine.detailed_mc(gens, base, lfu, 48, seed=5, year0=16)
    ryl, rhf = O.detailed_mc_philox(*_arrays(gens), base, lfu, 5, 16, 48)
    assert np.array_equal(yl.astype(float), ryl) and np.array_equal(hf.astype(float), rhf)
    a, _, _ = engine.detailed_mc(gens, base, lfu, 16, seed=5, year0=16)
    b, _, _ = engine.detailed_mc(gens, base, lfu, 32, seed=5, year0=32)
    assert np.array_equal(yl, np.concatenate([a, b]))


def test_run_detailed_mc_entry_point(engine):
    """tail_risk.jl:165 call shape: run_detailed_mc(gens, base_load, 5.0, 2000)."""
    gens, base, _ = _system()
    dist, prof = P.run_detailed_mc(gens, base, 5.0, 2000, seed=11, engine=engine)
    assert dist.shape == (2000,) and prof.shape == (8760,)
    assert abs(dist.mean() - prof.sum()) < 1e-9                  # both count the same deficit hours
    mean, prof2, yl = P.run_monte_carlo(gens, base, P.SystemParams(20.0, 5.0, 2000), seed=11, engine=engine)
    assert mean == dist.mean() and np.array_equal(prof, prof2)
    # no-ELU, no-maintenance, sigma = 0 limit equals the plain hourly Bernoulli model: compare with COPT
    plain = [P.DetailedGenerator(g.name, g.capacity, g.for_rate, 0) for g in gens]
    d0, p0 = P.run_detailed_mc(plain, base, 0.0, 4000, seed=3, engine=engine)
    lole, _, _ = O.analytical([g.capacity for g in plain], [g.for_rate for g in plain], base, 1.0)
    se = d0.std(ddof=1) / math.sqrt(len(d0))
    assert abs(d0.mean() - lole) < 4 * se + 0.05


def test_detailed_analytical_vs_literal_loops(engine):
    """tail_risk.jl:96-141: ELU effective-FOR fixed point + weekly COPT risk with 7-step LFU; the numeric cores
    (comprehensive.jl:118-142, tail_risk.jl:124-136) against the oracle's literal loops."""
    gens, base, _ = _system()
    lfu_mw = base.max() * 0.05
    rest = [g for g in gens if g.name != "Hydro_ELU"]
    probs = engine.copt([g.capacity for g in rest], [g.effective_q for g in rest], 20.0)
    e_fast = P.calculate_expected_generation(probs, 20.0, 200.0, base, lfu_mw)
    e_ref = O.expected_generation(probs, 20.0, 200.0, base, lfu_mw)
    assert abs(e_fast - e_ref) <= 1e-9 * e_ref and e_ref > 200.0 * 50.0        # the limit binds (comprehensive.jl:194-197)
    total, profile = P.run_detailed_analytical(gens, base, 5.0, engine=engine)
    hydro = [g for g in gens if g.name == "Hydro_ELU"][0]
    assert hydro.effective_q > hydro.for_rate + 1e-3                           # ELU raised the effective FOR
    assert abs(total - profile.sum()) < 1e-12 and (profile[8736:] == 0).all()
    for w in (1, 20, 52):                                                       # spot-check weeks against literal loops
        week = [g for g in gens if not (w >= g.scheduled_outage_start and w < g.scheduled_outage_start + g.maintenance_weeks)]
        pw = engine.copt([g.capacity for g in week], [g.effective_q for g in week], 20.0)
        ref = O.lfu_hourly_risk(pw, 20.0, base[(w - 1) * 168:w * 168], lfu_mw)
        assert np.allclose(profile[(w - 1) * 168:w * 168], ref, rtol=1e-9, atol=1e-15)
    # the MC mean sits above the analytical prediction (the point of tail_risk.jl plot 1)
    dist, _ = P.run_detailed_mc(gens, base, 5.0, 2000, seed=4, engine=engine)
    assert dist.mean() > 0.5 * total
