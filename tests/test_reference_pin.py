"""Pins the CPU oracle to the reference's own source text (CPU only).

tests/golden/ref_*.npz were produced by executing run_sequential_mc / run_non_sequential_mc / add_unit_convolution /
run_analytical of GeneratingAdequacy/PowerSystemAdequacy.jl, cut out of the reference checkout and transliterated line by
line into Python (oracle/jl_transliterate.py; generator: scripts/make_reference_golden.py).  Here
  * the C oracle (oracle/psra_oracle.c) and the independent hand transcription (oracle/psa_literal.py) must reproduce
    those vectors bit for bit -- everywhere, also on the GPU box where /root/reference does not exist;
  * when /root/reference exists, the vectors are re-derived from its text and must equal the committed ones, and the
    three draws that tools/patched_reference.jl / the transliteration substitute sit on PSA.jl:224,243,246, once each.
The CUDA path is held to the same vectors in tests/test_gpu_golden.py."""
import hashlib
import os
import re

import numpy as np
import pytest

from oracle import jl_transliterate as J
from oracle import m_transliterate as MT
from oracle import oracle as O
from oracle import psa_literal as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"
have_ref = os.path.exists(os.path.join(REF, J.PSA_REL))
SEQ_CASES = ("rts79_int", "rts79_mw", "small")


def _g(name):
    return np.load(os.path.join(G, f"ref_{name}.npz"))


def _history(lole, every):
    c = np.cumsum(lole)
    k = np.arange(every, len(lole) + 1, every)
    return c[k - 1] / k


@pytest.mark.parametrize("case", SEQ_CASES)
def test_c_oracle_reproduces_the_reference_sequential_loop(case):
    g = _g("seq_" + case)
    years = len(g["lole"])
    lol, eue, ent, used = O.seq_literal(g["cap"], g["load"], years, g["dur"])
    assert np.array_equal(lol, g["lole"]) and np.array_equal(eue, g["eue"])          # Float64, bit for bit
    assert lol.sum() > 0 and used.max() < g["dur"].shape[1]
    assert np.array_equal(_history(lol, 10), g["history"])                           # PSA.jl:264-266
    assert lol.sum() / years == float(g["lole_hours_yr"])
    tot = 0.0
    for e in eue:
        tot += e                                                                     # cum_eue += year_eue, in order
    assert tot / years == float(g["eue_mwh_yr"])


@pytest.mark.parametrize("case", SEQ_CASES)
def test_hand_transcription_reproduces_the_reference_sequential_loop(case):
    g = _g("seq_" + case)
    years = len(g["lole"])
    lole, eue, nlc, hist = L.sequential_mc(list(g["cap"]), list(g["load"]), years, [list(r) for r in g["dur"]])
    assert lole == list(g["lole"]) and eue == list(g["eue"]) and hist == list(g["history"])
    # deficit entries (calnlc.m) of the transcription == the C oracle's count
    _, _, ent, _ = O.seq_literal(g["cap"], g["load"], years, g["dur"])
    assert nlc == list(ent.astype(int))


def test_c_oracle_and_transcription_reproduce_the_reference_non_sequential_loop():
    g = _g("nonseq_rts79_mw")
    q = (1.0 / g["mttf"]) / ((1.0 / g["mttf"]) + (1.0 / g["mttr"]))                  # PSA.jl:32-37
    assert np.array_equal(q, g["for_rate"])
    lol, eue, _ = O.nonseq_literal(g["cap"], q, g["load"], g["r"])
    assert np.array_equal(lol, g["lole"]) and np.array_equal(eue, g["eue"])
    assert np.array_equal(_history(lol, 100), g["history"]) and lol.sum() / len(lol) == float(g["lole_hours_yr"])
    l2, e2, h2 = L.non_sequential_mc(list(g["cap"]), list(q), list(g["load"]), [list(r) for r in g["r"]])
    assert l2 == list(g["lole"]) and e2 == list(g["eue"]) and h2 == list(g["history"])


@pytest.mark.parametrize("step", [10, 7])
def test_c_oracle_reproduces_the_reference_analytical_engine(step):
    g = _g(f"analytical_step{step}")
    lole, eue, probs = O.analytical(g["cap"], g["for_rate"], g["load"], float(g["step"]))
    assert np.array_equal(probs, g["probs"])                                         # COPT, bit for bit
    # the transliteration sums sequentially where Julia sums pairwise: the indices agree to rounding, bar 1e-9
    assert abs(lole - float(g["lole"])) <= 1e-12 * lole and abs(eue - float(g["eue"])) <= 1e-12 * eue


def test_random_systems_oracle_vs_transcription():
    """20 seeded random systems: the C oracle and the hand transcription agree on every year (injected durations and
    injected uniforms) -- two restatements written independently of each other."""
    rng = np.random.default_rng(77)
    for _ in range(20):
        U = int(rng.integers(1, 12)); H = int(rng.integers(5, 200)); years = int(rng.integers(1, 25))
        cap = rng.integers(1, 80, U).astype(np.float64)
        mttf = rng.uniform(3.0, 300.0, U); mttr = rng.uniform(0.5, 40.0, U)
        load = np.round(rng.uniform(0.3, 1.0, H) * cap.sum(), int(rng.integers(0, 3)))
        K = int(4 * years * H / 3.0) + 16
        dur = np.empty((U, K))
        for u in range(U):
            dur[u, 0::2] = rng.exponential(mttf[u], len(dur[u, 0::2]))
            dur[u, 1::2] = rng.exponential(mttr[u], len(dur[u, 1::2]))
        dur[:, 3::4] *= rng.choice([1.0, 1e-2])
        dur = np.maximum(dur, 1e-9)
        lol, eue, ent, _ = O.seq_literal(cap, load, years, dur)
        a, b, c, _ = L.sequential_mc(list(cap), list(load), years, [list(r) for r in dur])
        assert a == list(lol) and b == list(eue) and c == list(ent.astype(int))
        q = (1.0 / mttf) / ((1.0 / mttf) + (1.0 / mttr))
        r = rng.random((40, U))
        nl, ne, _ = O.nonseq_literal(cap, q, load, r)
        a, b, _ = L.non_sequential_mc(list(cap), list(q), list(load), [list(x) for x in r])
        assert a == list(nl) and b == list(ne)


def test_calnlc_and_matlab_sampling_transcriptions():
    assert L.calnlc([1, 1, 0, 1, 0, 0, 1]) == 3 and L.calnlc([0, 0, 0]) == 0 and L.calnlc([0, 1, 1, 1]) == 1   # calnlc.m:22-34
    # seq_mcsampling.m:40-74: -450 ln(0.5) = 311.9 -> UP for 312 h; -50 ln(0.9) = 5.27 -> DOWN for 6 h from hour 313
    s, used = L.matlab_unit_series(450.0, 50.0, 400, [0.5, 0.9, 0.01, 0.3])
    assert sum(s) == 6 and s[312:318] == [1] * 6 and s[311] == 0 and used == 3


MATLAB_CASES = ("rts79", "small")


@pytest.mark.parametrize("name", MATLAB_CASES)
def test_c_oracle_reproduces_the_reference_matlab_sampler_and_calnlc(name):
    """ref_matlab.npz: seq_mcsampling.m (round / ceil next-event sampler, durations = the sampler streams) evaluated at HL1 and
    calnlc.m, both executed from the reference text (oracle/m_transliterate.py): DLC, ENS and NLC of every year."""
    g = _g("matlab")
    lol, eue, ent = O.seq_matlab_philox(g[f"{name}_cap"], g[f"{name}_mttf"], g[f"{name}_mttr"], g[f"{name}_load"],
                                        int(g[f"{name}_seed"]), int(g[f"{name}_year0"]), int(g[f"{name}_years"]))
    assert np.array_equal(lol, g[f"{name}_lol"]) and np.array_equal(eue, g[f"{name}_ens"]) and np.array_equal(ent, g[f"{name}_nlc"])
    assert lol.sum() > 100 and ent.sum() > 30


# ------------------------------------------------------------------------------------ needs the reference checkout
@pytest.mark.skipif(not have_ref, reason="/root/reference is not present (GPU box)")
def test_substitutions_hit_exactly_the_three_draws_of_the_reference():
    src = J.load_reference(REF)
    _, hit = J.apply_substitutions(src, J.SEQ_DRAW_SUBSTITUTIONS)
    assert hit == [224, 243, 246]
    # tools/patched_reference.jl performs the same three replacements on a real Julia
    tool = open(os.path.join(ROOT, "tools", "patched_reference.jl"), encoding="utf-8").read()
    olds = re.findall(r'must_replace\(src, "([^"]+)"', tool)
    assert olds[:3] == [s[1] for s in J.SEQ_DRAW_SUBSTITUTIONS]
    lines = src.split("\n")
    for (line_no, old, _) in J.SEQ_DRAW_SUBSTITUTIONS:
        assert src.count(old) == 1 and old in lines[line_no - 1]
    assert olds[3] == J.SEQ_RECORD_SUBSTITUTION[0] and src.count(olds[3]) == 1


@pytest.mark.skipif(not have_ref, reason="/root/reference is not present (GPU box)")
def test_committed_vectors_are_what_the_reference_text_produces():
    src = J.load_reference(REF)
    sha = hashlib.sha256(src.encode("utf-8")).hexdigest()
    g = _g("seq_small")
    assert str(g["reference_sha256"]) == sha
    res, lole, eue, hit = J.reference_sequential(src, g["cap"], g["mttf"], g["mttr"], g["load"], len(g["lole"]), g["dur"])
    assert hit == [224, 243, 246] and lole == list(g["lole"]) and eue == list(g["eue"])
    assert res.convergence_history == list(g["history"]) and res.method == "Sequential MC"
    g = _g("nonseq_rts79_mw")
    res, lole, eue, q = J.reference_non_sequential(src, g["cap"], g["mttf"], g["mttr"], g["load"], 100, g["r"][:100].reshape(-1))
    assert lole == list(g["lole"][:100]) and eue == list(g["eue"][:100]) and q == list(g["for_rate"])
    g = _g("analytical_step10")
    res, probs, _ = J.reference_analytical(src, g["cap"], g["mttf"], g["mttr"], g["load"], 10.0)
    assert probs == list(g["probs"]) and res.lole_hours_yr == float(g["lole"]) and res.eue_mwh_yr == float(g["eue"])
    # the known answers of SURVEY.md 8c / BASELINE.md section 3 come out of the reference text
    assert abs(res.lole_hours_yr - 9.420474608) < 1e-8 and abs(res.eue_mwh_yr - 1177.243237) < 1e-5


# ============================================================ the "next" rows: AdequacyAssessmentII.jl, tail_risk.jl, comprehensive.jl
def test_c_oracle_reproduces_the_reference_max_flow_curtailment():
    """solve_curtailment_fast (AdequacyAssessmentII.jl:73-179, transliterated from the reference text) on 400 random
    margin / topology cases, both policies: the C oracle returns the same curtailments, bit for bit."""
    g = _g("solve_curtailment")
    for topo, m, pol, cur, n in zip(g["topology"], g["margins"], g["policy"], g["curtailment"], g["n_areas"]):
        n = int(n)
        got = O.solve_curtailment(np.ascontiguousarray(topo[:n, :n]), m[:n], int(pol))
        assert np.array_equal(got, cur[:n]), (topo[:n, :n], m[:n], pol)
    assert (g["curtailment"] > 0).any() and (g["policy"] == 1).any()


@pytest.mark.parametrize("name", ["demo2", "mesh3"])
def test_c_oracle_reproduces_the_reference_multi_area_years(name):
    """run_fast_sequential_simulation (AdequacyAssessmentII.jl:185-250, transliterated; its two duration draws at :210,213 and
    the constructor's at :25 injected with the sampler durations of the (seed; year, unit) streams): hours with curtailment
    and curtailed energy per area and year == the C oracle driven by the same streams, both policies."""
    g = _g("multi_area_" + name)
    ny = g["lole"].shape[1]
    for pol in (0, 1):
        lol, eue = O.multi_area_philox(g["unit_area"], g["cap"], g["mttf"], g["mttr"], g["loads"], g["topology"], pol,
                                       int(g["seed"]), int(g["year0"]), ny, init_mode=0)
        assert np.array_equal(lol, g["lole"][pol]) and np.array_equal(eue, g["eue"][pol])
    assert g["lole"].sum() > 0 and not np.array_equal(g["eue"][0], g["eue"][1])      # the ties do carry support


def test_c_oracle_reproduces_the_reference_detailed_mc_and_maintenance_schedule():
    """schedule_maintenance! (generating_adequacy_comprehensive.jl:86-112) and run_detailed_mc (tail_risk.jl:12-91), both
    transliterated from the reference text; rand() / randn() replayed in the reference's own consumption order (no draw for
    a unit on maintenance).  The oracle's injected layout is per (year, hour, unit) -- the same values reach the same
    units, so the yearly LOLE counts and the hourly failure probabilities must be identical."""
    import powersystemsreliabilityassessment_b200.api as A
    g = _g("detailed_mc")
    assert O.schedule_maintenance(g["cap"], g["maint_weeks"], g["weekly_peaks"]) == list(g["maint_start"])
    gens = [A.DetailedGenerator(f"g{i}", float(c), float(q), int(w), float(e))
            for i, (c, q, w, e) in enumerate(zip(g["cap"], g["for_rate"], g["maint_weeks"], g["energy_limit"]))]
    A.schedule_maintenance(gens, list(g["weekly_peaks"]))                            # the product's host function
    assert [x.scheduled_outage_start for x in gens] == list(g["maint_start"])
    lfu_std = float(g["base_load"].max()) * (float(g["lfu_sigma_percent"]) / 100.0)   # tail_risk.jl:22
    yl, hf = O.detailed_mc_injected(g["cap"], g["for_rate"], g["maint_start"], g["maint_weeks"], g["energy_limit"], g["base_load"],
                                    lfu_std, g["unif"], g["norm"])
    assert np.array_equal(yl, g["yearly_lole"]) and yl.sum() > 0
    assert np.array_equal(hf / len(yl), g["hourly_failure_prob"])


@pytest.mark.skipif(not have_ref, reason="/root/reference is not present (GPU box)")
def test_next_row_vectors_are_what_the_reference_text_produces():
    src_ma = J.load_text(REF, J.MULTI_AREA_REL)
    _, hit = J.apply_substitutions(src_ma, J.MULTI_AREA_DRAW_SUBSTITUTIONS)
    assert hit == [210, 213]
    g = _g("solve_curtailment")
    for k in range(0, 400, 7):
        n = int(g["n_areas"][k])
        got = J.reference_solve_curtailment(src_ma, g["topology"][k][:n, :n], g["margins"][k][:n], int(g["policy"][k]))
        assert got == list(g["curtailment"][k][:n])
    g = _g("detailed_mc")
    assert J.reference_schedule_maintenance(J.load_text(REF, J.COMPREHENSIVE_REL), g["cap"], g["maint_weeks"], g["weekly_peaks"]) == list(g["maint_start"])
    yl, hf = J.reference_detailed_mc(J.load_text(REF, J.TAIL_RISK_REL), g["cap"], g["for_rate"], g["maint_start"], g["maint_weeks"],
                                     g["energy_limit"], g["base_load"], float(g["lfu_sigma_percent"]), 1, g["unif"][:1], g["norm"][:1])
    assert yl == [float(g["yearly_lole"][0])]


def test_c_oracle_reproduces_the_reference_fd_recursion_and_copt_demo():
    """generating_adequacy_frequency.jl (GeneratorFD constructor, add_unit_educational!, evaluate_risk) and
    generating_adequacy_assessment.jl (add_unit, calculate_indices), transliterated from the reference text: cumulative
    probability / frequency tables bit for bit, the risk indices and the COPT indices; the file's own demo gives the
    346.9045 / 3.8416 / 90.3022 of SURVEY.md 8c."""
    g = _g("fd")
    for k in range(int(g["n"])):
        P, F = O.fd_build(g[f"cap{k}"], g[f"mtbf{k}"], g[f"mttr{k}"])
        assert np.array_equal(P, g[f"P{k}"]) and np.array_equal(F, g[f"F{k}"])
        risk = O.fd_evaluate(P, F, float(g[f"peak{k}"]), float(g[f"cap{k}"].sum()))
        assert tuple(risk) == tuple(g[f"risk{k}"])
    assert abs(g["risk0"][0] - 346.9045) < 5e-5 and abs(g["risk0"][1] - 3.8416) < 5e-5 and abs(g["risk0"][2] - 90.3022) < 5e-5
    g = _g("gaa")
    for k in range(int(g["n"])):
        probs = O.gaa_build(g[f"cap{k}"], g[f"q{k}"], float(g[f"step{k}"]))
        assert np.array_equal(probs, g[f"probs{k}"])
        lole, eue = O.gaa_indices(probs, float(g[f"step{k}"]), g[f"ldc{k}"])
        assert abs(lole - g[f"idx{k}"][0]) <= 1e-12 * lole and abs(eue - g[f"idx{k}"][1]) <= 1e-12 * eue


def test_c_oracle_reproduces_the_reference_markov_script_blocks():
    """Markov_process.jl is a script; its constant-hazard experiment (:46-60), its two-state chain (:83-110) and its
    five-generator hourly DTMC (:153-195, the script's own unit data) are cut out as they stand, transliterated and run (on
    recorded rand() values where they draw)."""
    g = _g("markov")
    assert np.array_equal(O.dtmc_capacity(g["mttf"], g["mttr"], g["cap"], g["dtmc_uniforms"]), g["dtmc_capacity"])
    o = O.failure_times(float(g["ft_lambda"]), 1.0, 5000.0, len(g["ft_uniforms"]), uniforms=g["ft_uniforms"])
    assert np.array_equal(o[o >= 0], g["failure_times"]) and 0 < len(g["failure_times"]) < len(g["ft_uniforms"])
    # PART 4 (:83-110): the two-state chain, probability of DOWN after 1..steps hours (script parameters 1000 / 50 h, 200 steps)
    for key in ("markov2_script", "markov2_b"):
        mttf, mttr, dt = g[key + "_params"]
        assert np.array_equal(O.markov2(float(mttf), float(mttr), float(dt), len(g[key])), g[key])
    assert abs(g["markov2_script"][-1] - 0.047333337093) < 1e-12 and list(g["markov2_script_params"]) == [1000.0, 50.0, 1.0]
    if have_ref:
        src = J.load_text(REF, J.MARKOV_REL)
        series, mf, mr, cp = J.reference_dtmc_capacity(src, g["dtmc_uniforms"])
        assert series == list(g["dtmc_capacity"]) and mf == list(g["mttf"]) and cp == list(g["cap"])
        pd, lam, mu, dt, hit = J.reference_markov2(src)
        assert hit == [94, 102, 107] and pd == list(g["markov2_script"]) and (1 / lam, 1 / mu, dt) == (1000.0, 50.0, 1.0)
        assert J.reference_markov2(src, 450.0, 20.0, 500)[0] == list(g["markov2_b"])


def test_oracle_pieces_reproduce_the_reference_detailed_analytical():
    """run_detailed_analytical (tail_risk.jl:96-141) with update_elu! / calculate_expected_generation / add_unit of
    generating_adequacy_comprehensive.jl, transliterated and run on the six-unit system: the oracle's literal cores
    (expected_generation, lfu_hourly_risk, copt_build) composed the same way give the ELU fixed point and the hourly risk."""
    g = _g("detailed_mc"); r = _g("detailed_analytical")
    cap, q0, ms, mw, el, base = g["cap"], g["for_rate"], g["maint_start"], g["maint_weeks"], g["energy_limit"], g["base_load"]
    U = len(cap)
    lfu_mw = float(base.max()) * (5.0 / 100.0)
    q = q0.copy()
    hist = [float(q0[4])]
    for _ in range(5):                                                     # tail_risk.jl:106
        for i in range(U):
            if el[i] == np.inf:
                continue
            rest = [j for j in range(U) if j != i]
            probs = O.copt_build(cap[rest], q[rest], 20.0)
            req = O.expected_generation(probs, 20.0, float(cap[i]), base, lfu_mw)
            new_q = float(q0[i])
            if req > el[i]:
                new_q += (req - el[i]) / (cap[i] * len(base))
            new_q = min(new_q, 1.0)
            if abs(new_q - q[i]) > 1e-5:
                q[i] = new_q
            hist.append(new_q)
    assert np.allclose(hist, r["history_q_elu"], rtol=1e-12, atol=0) and np.allclose(q, r["effective_q"], rtol=1e-12, atol=0)
    assert r["effective_q"][4] > q0[4] + 1e-3                              # the energy limit binds
    prof = np.zeros(len(base))
    for w in range(1, 53):
        week = [j for j in range(U) if not (w >= ms[j] and w < ms[j] + mw[j])]
        probs = O.copt_build(cap[week], r["effective_q"][week], 20.0)
        h0, h1 = (w - 1) * 168, min(w * 168, 8760)
        prof[h0:h1] = O.lfu_hourly_risk(probs, 20.0, base[h0:h1], lfu_mw)
    assert np.allclose(prof, r["profile"], rtol=1e-12, atol=1e-300) and abs(prof.sum() - float(r["total"])) <= 1e-12 * prof.sum()


@pytest.mark.skipif(not have_ref, reason="/root/reference is not present (GPU box)")
def test_matlab_vectors_are_what_the_reference_text_produces():
    """seq_mcsampling.m / calnlc.m re-transliterated from the checkout: the two substituted draws sit on :52,59 (once each), the
    committed vectors come out again (RTS-79: the first 3 years; the small system: all 40), and the hand transcriptions of
    oracle/psa_literal.py agree with the reference's calnlc on random series."""
    sampler, _, hit = MT.load_seq_mcsampling(REF)
    calnlc, _ = MT.load_calnlc(REF)
    assert hit == [52, 59]
    g = _g("matlab")
    sha = hashlib.sha256((MT._load(REF, MT.SAMPLING_REL) + MT._load(REF, MT.CALNLC_REL)).encode("utf-8")).hexdigest()
    assert sha == str(g["reference_sha256"])
    for name, years, K in (("rts79", 3, 96), ("small", 40, 400)):
        cap, load = g[f"{name}_cap"], g[f"{name}_load"]
        mf = g[f"{name}_mttf"].astype(np.float32).astype(np.float64); mr = g[f"{name}_mttr"].astype(np.float32).astype(np.float64)
        seed, year0 = int(g[f"{name}_seed"]), int(g[f"{name}_year0"])
        for y in range(years):
            D = np.empty((len(cap), K))
            for u in range(len(cap)):
                words = []
                for b in range((K + 1 + 3) // 4):
                    words.extend(int(w) for w in O.philox([(year0 + y) & 0xffffffff, (year0 + y) >> 32, u, b], [seed & 0xffffffff, seed >> 32]))
                for k in range(K):
                    D[u, k] = O.duration_hours(mf[u] if k % 2 == 0 else mr[u], words[k + 1])
            states, _ = sampler(np.stack([mf, mr], axis=1), len(cap), 0, 1, len(load), D)
            st = np.array(states)
            cap_avail = ((1.0 - st) * cap[:, None]).sum(axis=0)
            flag = cap_avail < load
            assert flag.sum() == g[f"{name}_lol"][y] and (load - cap_avail)[flag].sum() == g[f"{name}_ens"][y]
            assert calnlc(flag.astype(np.float64)) == g[f"{name}_nlc"][y]
            assert np.array_equal(st.sum(axis=1), g[f"{name}_down_hours"][y])
    rng = np.random.default_rng(3)
    for _ in range(200):
        s = (rng.random(int(rng.integers(1, 60))) < rng.random()).astype(int).tolist()
        assert L.calnlc(s) == calnlc([float(x) for x in s])
