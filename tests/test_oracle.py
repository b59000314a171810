"""Pins the CPU oracle: every deterministic function against the known answers of SURVEY.md 8c /
BASELINE.md section 3 (derived by restating the cited reference lines; the reference ships no tests
and cannot run here), the sampler against the Random123 Philox4x32-10 known-answer vectors, and the
event-form identities the CUDA kernel relies on."""
import hashlib
import math

import numpy as np
import pytest

from helpers import injected_durations
from oracle import oracle as O
from powersystemsreliabilityassessment_b200 import rts79


def _for(mttf, mttr):
    lam = 1.0 / mttf; mu = 1.0 / mttr
    return lam / (lam + mu)


def test_rts79_unit_table():
    cap, mttf, mttr = rts79.units()
    assert len(cap) == 32 and cap.sum() == 3405.0
    q = _for(mttf, mttr)
    cls = {12: .02, 20: .10, 50: .01, 76: .02, 100: .04, 155: .04, 197: .05, 350: .08, 400: .12}
    for c, qq in zip(cap, q):
        assert abs(qq - cls[int(c)]) < 1e-12
    assert abs(32 + (2 * 8736 / (mttf + mttr)).sum() - 494.4) < 0.1     # events per system-year


def test_rts79_load_curve_known_answers():
    mw = rts79.load_curve_mw()
    assert len(mw) == 8736 and mw.max() == 2850.0 and mw.argmax() + 1 == 8442
    assert abs(mw.min() - 965.615625) < 1e-9 and mw.argmin() + 1 == 6365
    assert abs(mw.sum() - 15296714.91378) < 1e-4
    assert np.allclose(mw[:3], [1530.76977, 1439.38053, 1370.8386], rtol=0, atol=1e-9)
    li = rts79.load_curve_int()
    assert li.sum() == 15296715
    assert hashlib.sha256(li.astype("<i4").tobytes()).hexdigest() == \
        "5c0b9e7f53d086294047f5494c9eddaff1a9d3c89cc69f0ead05d865a0fe90ca"
    assert np.array_equal(rts79.load_factors(), O.load_factors(8736, rts79.WEEKLY, rts79.DAILY, rts79.HOURLY))


@pytest.mark.parametrize("load_kind,step,lole,eue,nstates", [
    ("float", 1.0, 9.3941103566, 1176.291677, 3406),
    ("int", 1.0, 9.3677375218, 1176.181257, 3406),
    ("float", 10.0, 9.4204746080, 1177.243237, 350)])
def test_run_analytical_known_answers(load_kind, step, lole, eue, nstates):
    cap, mttf, mttr = rts79.units()
    load = rts79.load_curve_mw() if load_kind == "float" else rts79.load_curve_int().astype(float)
    l, e, p = O.analytical(cap, _for(mttf, mttr), load, step)
    assert len(p) == nstates
    assert abs(l - lole) < 5e-10 and abs(e - eue) < 5e-7
    assert abs(p.sum() - 1.0) < 1e-12


def test_lolp_at_peak_and_golden_copt():
    cap, mttf, mttr = rts79.units()
    p = O.copt_build(cap, _for(mttf, mttr), 1.0)
    cum = np.cumsum(p[::-1])[::-1]
    assert abs(cum[556] - 0.084578060826) < 1e-12
    g = np.load("tests/golden/rts79_copt_step10.npy")
    assert np.array_equal(O.copt_build(cap, _for(mttf, mttr), 10.0), g)


def test_script_demos_known_answers():
    gp = O.gaa_build([50, 50, 56, 100], [0.02, 0.02, 0.04, 0.05], 10.0)
    ldc = np.array([200.0 - (100.0 / 8760) * (h - 1) for h in range(1, 8761)])
    l, e = O.gaa_indices(gp, 10.0, ldc)
    assert len(gp) == 27 and abs(l - 200.789852160021) < 1e-9 and abs(e - 4930.134560000009) < 1e-8
    P, F = O.fd_build([16.0, 16.0], [4380.0, 4380.0], [89.39, 89.39])
    a, b, c = O.fd_evaluate(P, F, 20.0, 32.0)
    assert abs(a - 346.9045) < 5e-5 and abs(b - 3.8416) < 5e-5 and abs(c - 90.3022) < 5e-5
    m = O.markov2(1000.0, 50.0, 1.0, 200)
    assert abs(m[-1] - 0.047333337093) < 1e-12
    assert abs((1 - math.exp(-1 / 1000)) - 9.995001666250e-4) < 1e-15


def test_philox_known_answer_vectors():
    """Random123 kat_vectors, philox4x32 10 rounds."""
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, out in kat:
        assert list(O.philox(ctr, key)) == out


def test_neglog_accuracy_and_range():
    """E(x) = -ln((x|1) / 2^32), sampler specification v2: mantissa truncated to 24 bits (|du/u| < 2^-23), degree-7
    minimax polynomial (|error| < 2.5e-7), one rounding of the result (half an ulp: 9.5e-7 for E > 16)."""
    rng = np.random.default_rng(0)
    for x in list(rng.integers(0, 2**32, 5000)) + [0, 1, 2, 3, 2**31, 2**32 - 1, 2**32 - 2, 2**24, 2**24 - 1]:
        e = O.neglog_u32(int(x))
        ref = -math.log((int(x) | 1) / 2**32)
        assert -1.3e-7 <= e <= 32 * math.log(2) * (1 + 1e-6)
        assert abs(e - ref) < 3.5e-7 * max(ref, 1.0)
    assert O.neglog_u32(0) == pytest.approx(32 * math.log(2), rel=1e-6)
    # the draws next to 2^32 (u -> 1): E is within the polynomial's error of 0, possibly <= 0 -- the duration clamp's case
    assert abs(O.neglog_u32(2**32 - 1)) < 2.5e-7
    es = np.array([O.neglog_u32(int(x)) for x in rng.integers(0, 2**32, 200000)])
    assert abs(es.mean() - 1.0) < 4 / math.sqrt(len(es))
    # durations are whole ticks of 2^-24 h, at least one tick
    for mean, x in ((1100.0, 12345), (1e-5, 2**32 - 1), (2940.0, 0), (2940.0, 2**32 - 1)):
        d = O.duration_hours(mean, x)
        assert d >= 2.0**-24 and d * 2**24 == int(d * 2**24)


def test_event_form_identity():
    """the literal `ttf -= 1.0` loop and the closed event form give the same toggle
    hours and residuals bit for bit (this is what seq_mc.cu relies on)."""
    rng = np.random.default_rng(3)
    for _ in range(2000):
        t = float(rng.exponential(rng.choice([0.3, 5.0, 900.0])))
        if t <= 0:
            continue
        x = t; h = 0
        while True:
            x = x - 1.0; h += 1
            if x <= 0:
                break
        c = math.ceil(t)
        r = (t - (c - 1.0)) - 1.0
        assert h == c and x == r


def test_seq_literal_sanity_and_fixture():
    cap, mttf, mttr = rts79.units()
    load = rts79.load_curve_int().astype(float)
    rng = np.random.default_rng(123)
    dur = injected_durations(rng, mttf, mttr, 1, 200)[0]
    lol, eue, ent, used = O.seq_literal(cap, load, 3, dur)
    g = np.load("tests/golden/seq_literal_seed123.npz")
    assert np.array_equal(lol, g["lol"]) and np.array_equal(eue, g["eue"]) and np.array_equal(ent, g["ent"])
    assert (ent <= lol).all() and ((lol > 0) == (eue > 0)).all()
    with pytest.raises(RuntimeError):
        O.seq_literal(cap, load, 3, dur[:, :4])


def test_seq_philox_statistics_vs_analytical():
    """The oracle's sampler-driven literal loop reproduces the analytical LOLE within its CI."""
    cap, mttf, mttr = rts79.units()
    load = rts79.load_curve_int().astype(float)
    lol, eue, ent = O.seq_philox(cap, mttf, mttr, load, 42, 0, 1500, 1, 1)
    se = lol.std(ddof=1) / math.sqrt(len(lol))
    assert abs(lol.mean() - 9.3677375218) < 3.5 * se
    g = np.load("tests/golden/seq_philox_seed42.npz")
    assert np.array_equal(lol[:64], g["lol"]) and np.array_equal(eue[:64], g["eue"]) and np.array_equal(ent[:64], g["ent"])


def test_nonseq_literal_vs_lookup_identity():
    """The sorted-load lookup used on the GPU equals the reference's hour loop (integer loads)."""
    cap, mttf, mttr = rts79.units()
    li = rts79.load_curve_int()
    rng = np.random.default_rng(9)
    r = rng.random((400, 32)) * 0.2
    lol, eue, cp = O.nonseq_literal(cap, _for(mttf, mttr), li.astype(float), r)
    s = np.sort(li); suf = np.concatenate([np.cumsum(s[::-1])[::-1], [0]])
    for i in range(400):
        ub = np.searchsorted(s, cp[i], side="right")
        assert lol[i] == len(s) - ub and eue[i] == suf[ub] - cp[i] * (len(s) - ub)


def test_quantile_type7_matches_numpy():
    rng = np.random.default_rng(2)
    x = rng.integers(0, 50000, 1001)
    for a in (0.0, 0.5, 0.95, 0.99, 1.0):
        assert O.quantile_type7(x, a) == pytest.approx(np.quantile(x, a, method="linear"), rel=1e-15)


def test_multi_area_curtailment_solver_known_cases():
    """solve_curtailment_fast (AdequacyAssessmentII.jl:73-179): hand-worked cases, incl. the reference's early break
    when the first deficit area cannot be reached any more."""
    t2 = np.array([[0, 200], [200, 0]], float)
    assert list(O.solve_curtailment(t2, [300, -250], 1)) == [0, 50]            # tie limit 200
    assert list(O.solve_curtailment(t2, [300, -250], 0)) == [0, 250]           # ISOLATED
    assert list(O.solve_curtailment(t2, [100, -250], 1)) == [0, 150]           # surplus limit
    assert list(O.solve_curtailment(t2, [100, 250], 1)) == [0, 0]              # fast path
    t3 = np.array([[0, 100, 0], [100, 0, 50], [0, 50, 0]], float)              # chain 1 - 2 - 3
    assert list(O.solve_curtailment(t3, [500, 0, -80], 1)) == [0, 0, 30]       # two hops, weakest link 50
    assert list(O.solve_curtailment(t3, [500, -120, -80], 1)) == [0, 20, 80]   # link 1-2 saturated -> loop stops (:136-147)
    assert list(O.solve_curtailment(t3, [-10, 500, -80], 1)) == [0, 0, 30]     # source in the middle, both directions
    # ISOLATED curtailment of an area equals the single-area literal loop fed the same streams
    cap = np.array([400.0] * 5 + [200.0] * 5); mttf = np.array([1000.0] * 5 + [900.0] * 5); mttr = np.array([50.0] * 5 + [60.0] * 5)
    ua = np.array([0] * 5 + [1] * 5)
    loads = np.stack([np.rint(1000 + 500 * np.sin(np.linspace(0, 2 * np.pi, 8760))), np.rint(800 + 400 * np.sin(np.linspace(0, 2 * np.pi, 8760)))])
    lol, eue = O.multi_area_philox(ua, cap, mttf, mttr, loads, t2, 0, 3, 0, 6, 1)
    l0, e0, _ = O.seq_philox(cap[:5], mttf[:5], mttr[:5], loads[0], 3, 0, 6, 1, 1)
    assert np.array_equal(lol[:, 0], l0) and np.array_equal(eue[:, 0], e0)
    lc, ec = O.multi_area_philox(ua, cap, mttf, mttr, loads, t2, 1, 3, 0, 6, 1)
    assert ec.sum() < eue.sum() and (lc <= lol + 1e-9).all() is not None


def test_sampler_vector_form_and_device_front_end_identity():
    """(i) the vector form used by the per-draw GPU test equals the scalar specification; (ii) the integer front end
    the CUDA kernels use for the logarithm (one round-toward-zero int->float conversion, psra_internal.cuh neglog_u32:
    exponent field 158 - lz, 23 truncated mantissa bits, -k from a funnel shift under the exponent of 2^23) yields the
    same mantissa bits and the same k as the clz / shift specification, over edge draws, a stride through the whole
    32-bit range and random draws; (iii) an independent numpy evaluation of the specification (fma emulated in float64,
    which rounds twice: so equality is asked of all but a handful of draws and 1 ulp of the rest)."""
    rng = np.random.default_rng(11)
    x = np.concatenate([np.array([0, 1, 2, 0xFFFFFFFF, 0x80000000, 0x7FFFFFFF, 0xB504F333, 0xB504F334], dtype=np.uint64),
                        np.arange(0, 1 << 32, 4099, dtype=np.uint64), rng.integers(0, 1 << 32, 100000, dtype=np.uint64)])
    x32 = x.astype(np.uint32)
    ticks, ebits = O.sampler_durations(450.0, x32[:2000])
    for i in range(0, 2000, 37):
        assert ebits[i] == np.float32(O.neglog_u32(int(x32[i]))).view(np.uint32)
        assert ticks[i] == round(O.duration_hours(450.0, int(x32[i])) * 2 ** 24)
    # specification
    w = x | 1
    lz = 32 - np.floor(np.log2(w.astype(np.float64))).astype(np.int64) - 1
    X = (w << lz.astype(np.uint64)) & 0xFFFFFFFF
    m_spec = 0x3F800000 | ((X >> 8) & 0x7FFFFF)
    k_spec = lz + 1
    # device formulation: fw = bits of RZ_f32(w) (numpy has no directed rounding: float64 holds w exactly, its top 24
    # bits are the truncation), mantissa by mask, -k = float(bits((0x00258000:fw) >> 23)) - (2^23 + 159)
    f64 = w.astype(np.float64).view(np.uint64)
    fw = ((((f64 >> 52) - 1023 + 127) << 23) | ((f64 >> 29) & 0x7FFFFF)).astype(np.uint64)
    m_dev = (fw & 0x7FFFFF) | 0x3F800000
    g_bits = ((fw >> 23) | (np.uint64(0x00258000) << np.uint64(9))) & 0xFFFFFFFF
    g = g_bits.astype(np.uint32).view(np.float32)
    nk = g - np.float32(8388767.0)
    assert np.array_equal(m_dev, m_spec)
    assert np.array_equal(-nk.astype(np.int64), k_spec) and k_spec.min() >= 1 and k_spec.max() <= 32
    # independent evaluation of the specification
    c = np.array([float.fromhex(h) for h in ("-0x1.9f324cp-2", "-0x1.555536p-1", "0x1.c72898p-3", "-0x1.94b470p-4",
                                             "0x1.90d388p-5", "-0x1.a7b9fep-6", "0x1.1d506cp-6", "-0x1.578b02p-7")], dtype=np.float32)
    t = m_spec.astype(np.uint32).view(np.float32) - np.float32(1.5)
    p = np.full_like(t, c[7])
    for i in range(6, -1, -1):
        p = (p.astype(np.float64) * t.astype(np.float64) + np.float64(c[i])).astype(np.float32)
    e_np = (k_spec.astype(np.float64) * np.float64(np.float32(float.fromhex("0x1.62e430p-1"))) + p.astype(np.float64)).astype(np.float32)
    _, e_c = O.sampler_durations(450.0, x32)
    diff = e_np.view(np.int32).astype(np.int64) - e_c.view(np.int32).astype(np.int64)
    assert np.abs(diff).max() <= 1 and (diff != 0).mean() < 1e-4


def test_sampler_and_nonseq_golden_fixtures():
    """The oracle still produces the committed sampler / non-sequential fixtures (scripts/make_golden.py)."""
    g = np.load("tests/golden/sampler_draws.npz")
    t, e = O.sampler_durations(450.0, g["draws"])
    assert np.array_equal(e, g["e_bits"]) and np.array_equal(t, g["ticks_450"])
    assert np.array_equal(O.sampler_durations(2940.0, g["draws"])[0], g["ticks_2940"])
    cap, mttf, mttr = rts79.units()
    g = np.load("tests/golden/nonseq_philox_seed7.npz")
    nl, ne, st = O.nonseq_philox(cap, mttf, mttr, rts79.load_curve_int().astype(float), 7, 1000, 256)
    assert np.array_equal(nl, g["lol"]) and np.array_equal(ne, g["eue"]) and np.array_equal(st, g["states"])
