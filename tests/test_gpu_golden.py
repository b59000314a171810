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


# ---- vectors produced by the reference's own source text (scripts/make_reference_golden.py, oracle/jl_transliterate.py)
def _ref(name):
    return np.load(os.path.join(G, f"ref_{name}.npz"))


@pytest.mark.parametrize("case", ["rts79_int", "small"])
def test_reference_text_sequential_vectors_integer_load(engine, case):
    """run_sequential_mc of PSA.jl:214-269 (transliterated from the reference text, the three draws injected) against
    psra_seq_eval_injected: every year's LOL hours and ENS, the history and the indices, bit for bit."""
    g = _ref("seq_" + case)
    years = len(g["lole"])
    engine.set_system(g["cap"], g["mttf"], g["mttr"]); engine.set_load(g["load"], strict=True)
    r = engine.seq_eval_injected(g["dur"][None, :, :], years_per_chain=years, group=10)
    assert np.array_equal(r.lol_hours.astype(np.float64), g["lole"]) and g["lole"].sum() > 0
    assert np.array_equal(r.raw["ens_fp_vector"].astype(np.float64), g["eue"])
    hist = np.cumsum(r.group_lol[:years // 10]) / (10.0 * np.arange(1, years // 10 + 1))
    assert np.array_equal(hist, g["history"])
    assert r.lole == float(g["lole_hours_yr"]) and r.eens == float(g["eue_mwh_yr"])


def test_reference_text_sequential_vectors_fractional_load(engine):
    """The RTS-79 curve in MW (fractional): the library puts the load on the integer grid with ceil, which keeps every
    loss-of-load hour of the Float64 comparison (c < L <=> c < ceil(L)); the deficit is over-stated by < 1 MW per hour."""
    g = _ref("seq_rts79_mw")
    years = len(g["lole"])
    engine.set_system(g["cap"], g["mttf"], g["mttr"]); engine.set_load(g["load"])
    r = engine.seq_eval_injected(g["dur"][None, :, :], years_per_chain=years)
    assert np.array_equal(r.lol_hours.astype(np.float64), g["lole"]) and g["lole"].sum() > 0
    over = r.raw["ens_fp_vector"].astype(np.float64) - g["eue"]
    assert (over >= 0).all() and (over < np.maximum(g["lole"], 1e-300)).all()


def test_reference_text_non_sequential_vectors(engine):
    g = _ref("nonseq_rts79_mw")
    engine.set_system(g["cap"], g["mttf"], g["mttr"]); engine.set_load(g["load"])
    r = engine.nonseq_eval_uniforms(g["r"], group=100)
    assert np.array_equal(r["lol_hours"].astype(np.float64), g["lole"])
    over = r["ens"].astype(np.float64) - g["eue"]
    assert (over >= -1e-9).all() and (over < np.maximum(g["lole"], 1e-300) + 1e-9).all()
    hist = np.cumsum(r["group_lol"]) / (100.0 * np.arange(1, 4))
    assert np.array_equal(hist, g["history"])


@pytest.mark.parametrize("step", [10, 7])
def test_reference_text_analytical_vectors(engine, step):
    g = _ref(f"analytical_step{step}")
    probs = engine.copt(g["cap"], g["for_rate"], float(g["step"]))
    assert np.array_equal(probs, g["probs"])                                  # bit-identical FP64 COPT
    lole, eue = engine.copt_indices(probs, float(g["step"]), float(g["cap"].sum()), g["load"])
    assert abs(lole - float(g["lole"])) <= 1e-9 * lole and abs(eue - float(g["eue"])) <= 1e-9 * eue


@pytest.mark.parametrize("name", ["demo2", "mesh3"])
def test_reference_text_multi_area_vectors(engine, name):
    """run_fast_sequential_simulation + solve_curtailment_fast of AdequacyAssessmentII.jl (transliterated from the reference
    text, durations = the sampler streams) against psra_multi_area_mc: per year and area, both policies."""
    import powersystemsreliabilityassessment_b200 as P
    g = _ref("multi_area_" + name)
    ny = g["lole"].shape[1]
    for pol in (P.ISOLATED, P.INTERCONNECTED):
        m = engine.multi_area_mc(g["unit_area"], g["cap"], g["mttf"], g["mttr"], g["loads"], g["topology"], pol, ny,
                                 seed=int(g["seed"]), year0=int(g["year0"]), init_mode=P.INIT_ALL_UP, per_year=True)
        assert np.array_equal(m["lol_hours"].astype(np.float64), g["lole"][pol])
        assert np.array_equal(m["ens_fp"].astype(np.float64), g["eue"][pol])


def test_reference_text_detailed_mc_vectors(engine):
    """run_detailed_mc of tail_risk.jl:12-91 (transliterated, recorded rand() / randn()) against psra_detailed_eval_injected."""
    import powersystemsreliabilityassessment_b200 as P
    g = _ref("detailed_mc")
    gens = [P.DetailedGenerator(f"g{i}", float(c), float(q), int(w), float(e))
            for i, (c, q, w, e) in enumerate(zip(g["cap"], g["for_rate"], g["maint_weeks"], g["energy_limit"]))]
    P.schedule_maintenance(gens, list(g["weekly_peaks"]))
    assert [x.scheduled_outage_start for x in gens] == list(g["maint_start"])
    lfu_std = float(g["base_load"].max()) * (float(g["lfu_sigma_percent"]) / 100.0)
    yl, hf = engine.detailed_eval_injected(gens, g["base_load"], lfu_std, g["unif"], g["norm"])
    assert np.array_equal(yl.astype(np.float64), g["yearly_lole"])
    assert np.array_equal(hf.astype(np.float64) / len(yl), g["hourly_failure_prob"])


def test_reference_text_fd_and_copt_demo_vectors(engine):
    """F&D recursion (generating_adequacy_frequency.jl) and the stand-alone COPT indices (generating_adequacy_assessment.jl),
    vectors from the transliterated reference text, against psra_fd_recursion / psra_copt_indices_strict."""
    from powersystemsreliabilityassessment_b200 import evaluate_risk
    g = _ref("fd")
    for k in range(int(g["n"])):
        P, F = engine.fd_recursion(g[f"cap{k}"], g[f"mtbf{k}"], g[f"mttr{k}"])
        assert np.array_equal(P, g[f"P{k}"]) and np.array_equal(F, g[f"F{k}"])           # FP64 tables, bit for bit
        assert tuple(evaluate_risk(P, F, float(g[f"peak{k}"]), float(g[f"cap{k}"].sum()))) == tuple(g[f"risk{k}"])
    g = _ref("gaa")
    for k in range(int(g["n"])):
        lole, eue = engine.copt_indices_strict(g[f"probs{k}"], float(g[f"step{k}"]), g[f"ldc{k}"])
        assert abs(lole - g[f"idx{k}"][0]) <= 1e-9 * lole and abs(eue - g[f"idx{k}"][1]) <= 1e-9 * eue


def test_reference_text_markov_script_vectors(engine):
    g = _ref("markov")
    assert np.array_equal(engine.dtmc_capacity(g["mttf"], g["mttr"], g["cap"], g["dtmc_uniforms"]), g["dtmc_capacity"])
    ft = engine.failure_times(float(g["ft_lambda"]), len(g["ft_uniforms"]), 1.0, 5000.0, uniforms=g["ft_uniforms"])
    assert np.array_equal(ft, g["failure_times"])


def test_reference_text_detailed_analytical_vectors(engine):
    """run_detailed_analytical of tail_risk.jl:96-141 (transliterated with the comprehensive.jl functions it calls) against the
    product's host functions over psra_copt: ELU effective FOR and the hourly risk profile within 1e-9."""
    import powersystemsreliabilityassessment_b200 as P
    g = _ref("detailed_mc"); r = _ref("detailed_analytical")
    gens = [P.DetailedGenerator(f"g{i}", float(c), float(q), int(w), float(e))
            for i, (c, q, w, e) in enumerate(zip(g["cap"], g["for_rate"], g["maint_weeks"], g["energy_limit"]))]
    for x, s in zip(gens, g["maint_start"]):
        x.scheduled_outage_start = int(s)
    total, profile = P.run_detailed_analytical(gens, g["base_load"], 5.0, engine=engine)
    assert np.allclose([x.effective_q for x in gens], r["effective_q"], rtol=1e-9, atol=0)
    assert np.allclose(profile, r["profile"], rtol=1e-9, atol=1e-300) and abs(total - float(r["total"])) <= 1e-9 * total


@pytest.mark.parametrize("name", ("rts79", "small"))
def test_matlab_sampler_vectors_from_the_reference_text(engine, name):
    """ref_matlab.npz: Montecarlo_seq/seq_mcsampling.m + calnlc.m executed from the reference text (oracle/m_transliterate.py) on
    the sampler's duration streams; the CUDA kernels in MATLAB-discretisation mode return the same DLC / ENS / NLC per year."""
    from powersystemsreliabilityassessment_b200 import DISC_MATLAB, INIT_ALL_UP, Engine
    g = _ref("matlab")
    cap, mttf, mttr, load = g[f"{name}_cap"], g[f"{name}_mttf"], g[f"{name}_mttr"], g[f"{name}_load"]
    seed, year0, years = int(g[f"{name}_seed"]), int(g[f"{name}_year0"]), int(g[f"{name}_years"])
    for kw in (dict(), dict(force_generic=True)):
        with Engine(**kw) as e:
            e.set_system(cap, mttf, mttr); e.set_load(load.astype(np.int32))
            r = e.seq_mc(years, seed=seed, year0=year0, init_mode=INIT_ALL_UP | DISC_MATLAB, per_year=True)
        assert np.array_equal(r.lol_hours.astype(np.float64), g[f"{name}_lol"])
        assert np.array_equal(r.raw["ens_fp_vector"].astype(np.float64), g[f"{name}_ens"])
        assert np.array_equal(r.entries.astype(np.float64), g[f"{name}_nlc"])


def test_two_state_markov_chain_from_the_reference_text(engine):
    """ref_markov.npz: Markov_process.jl:83-110 executed from the reference text; psra_markov2 returns the same series."""
    g = _ref("markov")
    for key in ("markov2_script", "markov2_b"):
        mttf, mttr, dt = g[key + "_params"]
        m = engine.markov2(float(mttf), float(mttr), float(dt), len(g[key]))
        assert np.allclose(m, g[key], rtol=0, atol=1e-15)
