"""GPU parity of psra_multi_area_mc (multi-area adequacy with tie-line support, SURVEY f-3) against the CPU oracle's
literal restatement of GeneratingAdequacy/AdequacyAssessmentII.jl:73-179,185-250 -- bit-exact per year and area."""
import numpy as np
import pytest

import powersystemsreliabilityassessment_b200 as P
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _demo(H=8760):
    """run_demo system (AdequacyAssessmentII.jl:256-270) with loads rounded to whole MW."""
    cap = np.array([400.0] * 5 + [200.0] * 5); mttf = np.array([1000.0] * 5 + [900.0] * 5); mttr = np.array([50.0] * 5 + [60.0] * 5)
    ua = np.array([0] * 5 + [1] * 5)
    l1 = np.rint(1000.0 + 500.0 * np.sin(np.linspace(0, 2 * np.pi, H)))
    l2 = np.rint(800.0 + 400.0 * np.sin(np.linspace(0, 2 * np.pi, H)))
    topo = np.array([[0.0, 200.0], [200.0, 0.0]])
    return ua, cap, mttf, mttr, np.stack([l1, l2]), topo


@pytest.mark.parametrize("policy", [P.ISOLATED, P.INTERCONNECTED])
@pytest.mark.parametrize("init_mode", [0, 1])
def test_demo_system_bit_exact(engine, policy, init_mode):
    ua, cap, mttf, mttr, loads, topo = _demo()
    r = engine.multi_area_mc(ua, cap, mttf, mttr, loads, topo, policy, 40, seed=9, year0=16, init_mode=init_mode, per_year=True)
    lol, eue = O.multi_area_philox(ua, cap, mttf, mttr, loads, topo, policy, 9, 16, 40, init_mode)
    assert np.array_equal(r["lol_hours"].astype(np.float64), lol) and lol.sum() > 0
    assert np.array_equal(r["ens_fp"].astype(np.float64), eue)
    assert np.array_equal(r["sum_lol_hours"], lol.sum(0).astype(np.int64)) and np.array_equal(r["sum_ens_fp"], eue.sum(0).astype(np.int64))
    # years are independent: any split gives the same integers
    a = engine.multi_area_mc(ua, cap, mttf, mttr, loads, topo, policy, 15, seed=9, year0=16, init_mode=init_mode, per_year=True)
    b = engine.multi_area_mc(ua, cap, mttf, mttr, loads, topo, policy, 25, seed=9, year0=31, init_mode=init_mode, per_year=True)
    assert np.array_equal(np.concatenate([a["lol_hours"], b["lol_hours"]]), r["lol_hours"])


def test_meshed_four_areas_and_ragged_hours(engine):
    """4 areas on a ring with one weak link, unequal unit counts (incl. an area without units), H not a multiple of 32:
    exercises multi-hop augmenting paths, reverse residuals and the reference's early break."""
    rng = np.random.default_rng(4)
    H = 1000
    ua = np.array([0] * 7 + [1] * 3 + [3] * 12)
    cap = np.concatenate([rng.choice([50, 100, 150], 7), rng.choice([80, 120], 3), rng.choice([20, 40, 60], 12)]).astype(float)
    mttf = rng.uniform(80, 400, len(cap)); mttr = rng.uniform(10, 60, len(cap))
    t = np.arange(H)
    loads = np.stack([np.rint(380 + 150 * np.sin(t / 37.0)), np.rint(230 + 90 * np.cos(t / 53.0)),
                      np.rint(60 + 30 * np.sin(t / 11.0)), np.rint(300 + 120 * np.sin(t / 71.0 + 1))]).clip(0)
    topo = np.zeros((4, 4))
    for i, j, c in ((0, 1, 60), (1, 2, 40), (2, 3, 25), (3, 0, 80), (0, 2, 15)):
        topo[i, j] += c; topo[j, i] += c
    for policy in (P.ISOLATED, P.INTERCONNECTED):
        r = engine.multi_area_mc(ua, cap, mttf, mttr, loads, topo, policy, 60, seed=21, per_year=True)
        lol, eue = O.multi_area_philox(ua, cap, mttf, mttr, loads, topo, policy, 21, 0, 60, 1)
        assert np.array_equal(r["lol_hours"].astype(np.float64), lol) and np.array_equal(r["ens_fp"].astype(np.float64), eue)
        assert (lol.sum(0) > 0).sum() >= 3
    iso = engine.multi_area_mc(ua, cap, mttf, mttr, loads, topo, P.ISOLATED, 60, seed=21)
    con = engine.multi_area_mc(ua, cap, mttf, mttr, loads, topo, P.INTERCONNECTED, 60, seed=21)
    assert con["sum_ens_fp"].sum() < iso["sum_ens_fp"].sum()          # support can only reduce the total curtailment


def test_isolated_area_equals_single_area_kernel(engine):
    """ISOLATED = the HL1 sequential kernel per area (SURVEY f-3): same streams when the area's units come first."""
    ua, cap, mttf, mttr, loads, topo = _demo(8736)
    r = engine.multi_area_mc(ua, cap, mttf, mttr, loads, topo, P.ISOLATED, 64, seed=5, per_year=True)
    engine.set_system(cap[:5], mttf[:5], mttr[:5]); engine.set_load(loads[0].astype(np.int32))
    s = engine.seq_mc(64, seed=5, per_year=True)
    assert np.array_equal(r["lol_hours"][:, 0], s.lol_hours) and np.array_equal(r["ens_fp"][:, 0], s.raw["ens_fp_vector"])


def test_reference_entry_point_and_errors(engine):
    gens1 = [P.AreaGenerator(f"G1_{i}", 400.0, 1000.0, 50.0) for i in range(1, 6)]
    gens2 = [P.AreaGenerator(f"G2_{i}", 200.0, 900.0, 60.0) for i in range(1, 6)]
    load1 = 1000.0 + 500.0 * np.sin(np.linspace(0, 2 * np.pi, 8760)); load2 = 800.0 + 400.0 * np.sin(np.linspace(0, 2 * np.pi, 8760))
    sysm = P.System([P.Area(1, "Area_Rich", gens1, load1), P.Area(2, "Area_Poor", gens2, load2)], [P.TieLine(1, 2, 200.0)])
    assert np.array_equal(sysm.topology_matrix, [[0, 200], [200, 0]])
    iso = P.run_fast_sequential_simulation(sysm, P.ISOLATED, 2000, engine=engine, verbose=False)
    con = P.run_fast_sequential_simulation(sysm, P.INTERCONNECTED, 2000, engine=engine, verbose=False)
    assert [r["area"] for r in iso] == ["Area_Rich", "Area_Poor"]
    assert con[1]["lole"] < 0.5 * iso[1]["lole"] and con[1]["eue"] < 0.5 * iso[1]["eue"]      # the poor area is supported
    assert 2000 < iso[1]["lole"] < 5000 and 20 < iso[0]["lole"] < 120
    with pytest.raises(P.PsraError):
        engine.multi_area_mc([0, 9], [1, 1], [10, 10], [1, 1], np.ones((2, 64)), np.zeros((2, 2)), 0, 4)
    with pytest.raises(P.PsraError):
        engine.multi_area_mc([0] * 3, [1] * 3, [10] * 3, [1] * 3, np.ones((7, 8760)), np.zeros((7, 7)), 1, 4)   # shared memory
