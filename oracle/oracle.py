"""ctypes front-end of the CPU oracle (oracle/psra_oracle.c) + small numpy restatements.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
Parity status: pinned by SURVEY.md section 8c known answers (tests/test_oracle.py) and by vectors the reference's
own source text produces when transliterated line by line (oracle/jl_transliterate.py, tests/golden/ref_*.npz,
tests/test_reference_pin.py); the reference's Monte Carlo random streams themselves are irreproducible (unseeded Julia RNG).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpsra_oracle.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "psra_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.oracle_seq_literal.restype = C.c_int
        L.oracle_seq_literal.argtypes = [C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_void_p,
                                         _dp, _dp, _dp, C.c_void_p]
        L.oracle_philox4x32_10.restype = None
        L.oracle_philox4x32_10.argtypes = [_u32p, _u32p, _u32p]
        L.oracle_neglog_u32.restype = C.c_float
        L.oracle_neglog_u32.argtypes = [C.c_uint32]
        L.oracle_duration_hours.restype = C.c_double
        L.oracle_duration_hours.argtypes = [C.c_float, C.c_uint32]
        L.oracle_sampler_durations.restype = None
        L.oracle_sampler_durations.argtypes = [C.c_float, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]
        L.oracle_seq_philox.restype = C.c_int
        L.oracle_seq_philox.argtypes = [C.c_int, _dp, _fp, _fp, _u32p, C.c_int, _dp, C.c_uint64,
                                        C.c_int64, C.c_int64, C.c_int, C.c_int, _dp, _dp, _dp]
        L.oracle_seq_philox_ex.restype = C.c_int
        L.oracle_seq_philox_ex.argtypes = [C.c_int, _dp, _fp, _fp, _u32p, C.c_int, _dp, C.c_uint64,
                                           C.c_int64, C.c_int64, C.c_int, C.c_int, _dp, _dp, _dp, _dp]
        L.oracle_solve_curtailment.restype = None
        L.oracle_solve_curtailment.argtypes = [C.c_int, _dp, _dp, C.c_int, _dp]
        L.oracle_multi_area_philox.restype = C.c_int
        L.oracle_multi_area_philox.argtypes = [C.c_int, C.c_int, _i32p, _dp, _fp, _fp, _u32p, C.c_int, _dp, _dp, C.c_int,
                                               C.c_uint64, C.c_int64, C.c_int64, C.c_int, _dp, _dp]
        L.oracle_failure_times.restype = C.c_int
        L.oracle_failure_times.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int64, C.c_uint64, C.c_void_p, C.c_int, _dp]
        L.oracle_nonseq_literal.restype = None
        L.oracle_nonseq_literal.argtypes = [C.c_int, _dp, _dp, C.c_int, _dp, C.c_int64, _dp, _dp, _dp, _dp]
        L.oracle_nonseq_states.restype = None
        L.oracle_nonseq_states.argtypes = [C.c_int, _dp, C.c_int, _dp, C.c_int64, _u32p, _dp, _dp]
        L.oracle_nonseq_philox.restype = None
        L.oracle_nonseq_philox.argtypes = [C.c_int, _dp, _u32p, C.c_int, _dp, C.c_uint64, C.c_int64,
                                           C.c_int64, _dp, _dp, _u32p]
        L.oracle_copt_next_len.restype = C.c_int
        L.oracle_copt_next_len.argtypes = [C.c_int, C.c_double, C.c_double]
        L.oracle_copt_build.restype = C.c_int
        L.oracle_copt_build.argtypes = [C.c_int, _dp, _dp, C.c_double, _dp, C.c_int]
        L.oracle_analytical_indices.restype = None
        L.oracle_analytical_indices.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_int, _dp,
                                                C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_gaa_add_unit.restype = None
        L.oracle_gaa_add_unit.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_double, _dp, C.c_int]
        L.oracle_gaa_calculate_indices.restype = None
        L.oracle_gaa_calculate_indices.argtypes = [_dp, C.c_int, C.c_double, C.c_int, _dp,
                                                   C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_fd_add_unit.restype = None
        L.oracle_fd_add_unit.argtypes = [_dp, _dp, C.c_int, C.c_double, C.c_double, C.c_double,
                                         C.c_double, _dp, _dp, C.c_int]
        L.oracle_fd_evaluate.restype = None
        L.oracle_fd_evaluate.argtypes = [_dp, _dp, C.c_int, C.c_double, C.c_double,
                                         C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_markov2.restype = None
        L.oracle_markov2.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, _dp]
        L.oracle_dtmc_capacity.restype = None
        L.oracle_dtmc_capacity.argtypes = [C.c_int, _dp, _dp, _dp, C.c_int, _dp, _dp]
        L.oracle_seq_matlab_philox.restype = C.c_int
        L.oracle_seq_matlab_philox.argtypes = [C.c_int, _dp, _fp, _fp, C.c_int, _dp, C.c_uint64, C.c_int64, C.c_int64,
                                               _dp, _dp, _dp]
        L.oracle_normal_u32x2.restype = C.c_float
        L.oracle_normal_u32x2.argtypes = [C.c_uint32, C.c_uint32]
        L.oracle_detailed_mc_injected.restype = C.c_int
        L.oracle_detailed_mc_injected.argtypes = [C.c_int, _dp, _dp, _i32p, _i32p, _dp, C.c_int, _dp, C.c_double,
                                                  C.c_int, _dp, _dp, _dp, _dp]
        L.oracle_detailed_mc_philox.restype = C.c_int
        L.oracle_detailed_mc_philox.argtypes = [C.c_int, _dp, _u32p, _i32p, _i32p, _dp, C.c_int, _dp, C.c_double,
                                                C.c_uint64, C.c_int64, C.c_int, _dp, _dp]
        L.oracle_expected_generation.restype = C.c_double
        L.oracle_expected_generation.argtypes = [_dp, C.c_int, C.c_double, C.c_double, _dp, C.c_int, C.c_double]
        L.oracle_lfu_hourly_risk.restype = None
        L.oracle_lfu_hourly_risk.argtypes = [_dp, C.c_int, C.c_double, _dp, C.c_int, C.c_double, _dp]
        L.oracle_load_factors.restype = None
        L.oracle_load_factors.argtypes = [C.c_int, _dp, _dp, _dp, _dp]
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# --------------------------------------------------------------------------- sequential MC
def seq_literal(cap, load, years, durations, init_status=None):
    """PSA.jl:214-269 with injected per-unit duration lists durations[U, K].
    Returns (lol[years], eue[years], entries[years], draws_used[U])."""
    cap = _d(cap); load = _d(load); dur = _d(durations)
    U, K = dur.shape
    lol = np.zeros(years); eue = np.zeros(years); ent = np.zeros(years)
    used = np.zeros(U, dtype=np.int32)
    st = None
    if init_status is not None:
        st_arr = np.ascontiguousarray(init_status, dtype=np.uint8)
        st = st_arr.ctypes.data_as(C.c_void_p)
    rc = lib().oracle_seq_literal(U, cap, len(load), load, years, dur, K, st, lol, eue, ent,
                                  used.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("oracle_seq_literal: injected durations exhausted")
    return lol, eue, ent, used


def for_threshold(mttf, mttr):
    """floor(FOR * 2^32) with FOR = MTTR/(MTTF+MTTR) (PSA.jl:32-37, failprob.m:23)."""
    mttf = _d(mttf); mttr = _d(mttr)
    lam = 1.0 / mttf; mu = 1.0 / mttr
    q = lam / (lam + mu)
    return np.minimum(np.floor(q * 4294967296.0), 4294967295.0).astype(np.uint32)


def seq_philox(cap, mttf, mttr, load, seed, chain0, nchains, years_per_chain=1, init_mode=1):
    cap = _d(cap); load = _d(load)
    mf = np.ascontiguousarray(mttf, dtype=np.float32); mr = np.ascontiguousarray(mttr, dtype=np.float32)
    thr = for_threshold(mttf, mttr)
    n = nchains * years_per_chain
    lol = np.zeros(n); eue = np.zeros(n); ent = np.zeros(n)
    lib().oracle_seq_philox(len(cap), cap, mf, mr, thr, len(load), load, seed, chain0, nchains,
                            years_per_chain, init_mode, lol, eue, ent)
    return lol, eue, ent


def seq_philox_importance(cap, mttf, mttr, load, seed, chain0, nchains, years_per_chain=1, init_mode=1):
    """oracle_seq_philox plus unit_down_in_loss[U]: hours with loss of load in which each unit is DOWN
    (Montecarlo_seq/seqMain.m:140-150,225-231 restricted to generators)."""
    cap = _d(cap); load = _d(load)
    mf = np.ascontiguousarray(mttf, dtype=np.float32); mr = np.ascontiguousarray(mttr, dtype=np.float32)
    thr = for_threshold(mttf, mttr)
    n = nchains * years_per_chain
    lol = np.zeros(n); eue = np.zeros(n); ent = np.zeros(n); imp = np.zeros(len(cap))
    lib().oracle_seq_philox_ex(len(cap), cap, mf, mr, thr, len(load), load, seed, chain0, nchains,
                               years_per_chain, init_mode, lol, eue, ent, imp)
    return lol, eue, ent, imp


def solve_curtailment(topology, margins, policy):
    """AdequacyAssessmentII.jl:73-179 (policy 0 = ISOLATED, 1 = INTERCONNECTED)."""
    topo = _d(topology); m = _d(margins)
    out = np.zeros(len(m))
    lib().oracle_solve_curtailment(len(m), topo, m, int(policy), out)
    return out


def multi_area_philox(unit_area, cap, mttf, mttr, loads, topology, policy, seed, year0, nyears, init_mode=1):
    """AdequacyAssessmentII.jl:185-250 driven by the sampler streams; loads[n_areas][H]; returns per-year, per-area
    (hours with curtailment, curtailed energy)."""
    ua = np.ascontiguousarray(unit_area, dtype=np.int32); cap = _d(cap); loads = _d(loads); topo = _d(topology)
    mf = np.ascontiguousarray(mttf, dtype=np.float32); mr = np.ascontiguousarray(mttr, dtype=np.float32)
    thr = for_threshold(mttf, mttr)
    A, H = loads.shape
    lol = np.zeros((nyears, A)); eue = np.zeros((nyears, A))
    rc = lib().oracle_multi_area_philox(A, len(cap), ua, cap, mf, mr, thr, H, loads, topo, int(policy), seed, year0,
                                        nyears, init_mode, lol, eue)
    if rc:
        raise RuntimeError("oracle_multi_area_philox failed")
    return lol, eue


def failure_times(lam, dt, max_time, n, seed=42, uniforms=None):
    """Markov_process.jl:39-60; returns the per-component array (-1 = outlived max_time)."""
    out = np.zeros(n)
    if uniforms is None:
        rc = lib().oracle_failure_times(lam, dt, max_time, n, seed, None, 0, out)
    else:
        r = _d(uniforms)
        rc = lib().oracle_failure_times(lam, dt, max_time, n, seed, r.ctypes.data_as(C.c_void_p), r.shape[1], out)
    if rc:
        raise RuntimeError("oracle_failure_times: injected uniforms exhausted")
    return out


def seq_matlab_philox(cap, mttf, mttr, load, seed, year0, nyears):
    """Montecarlo_seq/seq_mcsampling.m:40-74 discretisation (round / ceil, all UP every year) at HL1."""
    cap = _d(cap); load = _d(load)
    mf = np.ascontiguousarray(mttf, dtype=np.float32); mr = np.ascontiguousarray(mttr, dtype=np.float32)
    lol = np.zeros(nyears); eue = np.zeros(nyears); ent = np.zeros(nyears)
    lib().oracle_seq_matlab_philox(len(cap), cap, mf, mr, len(load), load, seed, year0, nyears, lol, eue, ent)
    return lol, eue, ent


def philox(ctr, key):
    out = np.zeros(4, dtype=np.uint32)
    lib().oracle_philox4x32_10(np.asarray(ctr, dtype=np.uint32), np.asarray(key, dtype=np.uint32), out)
    return out


def duration_hours(mean: float, x: int) -> float:
    return float(lib().oracle_duration_hours(float(np.float32(mean)), int(x)))


def sampler_durations(mean: float, draws):
    """(ticks, e_bits) of an array of 32-bit draws: the sampler specification of DESIGN.md 3.2, draw by draw."""
    x = np.ascontiguousarray(draws, dtype=np.uint32)
    ticks = np.zeros(x.size, dtype=np.uint64)
    ebits = np.zeros(x.size, dtype=np.uint32)
    lib().oracle_sampler_durations(float(np.float32(mean)), x.ctypes.data, x.size, ticks.ctypes.data, ebits.ctypes.data)
    return ticks, ebits


def neglog_u32(x: int) -> float:
    return float(lib().oracle_neglog_u32(int(x)))


# ----------------------------------------------------------------------- non-sequential MC
def nonseq_literal(cap, for_rate, load, r):
    cap = _d(cap); q = _d(for_rate); load = _d(load); r = _d(r)
    n = r.shape[0]
    lol = np.zeros(n); eue = np.zeros(n); cp = np.zeros(n)
    lib().oracle_nonseq_literal(len(cap), cap, q, len(load), load, n, r, lol, eue, cp)
    return lol, eue, cp


def nonseq_states(cap, load, states):
    cap = _d(cap); load = _d(load)
    st = np.ascontiguousarray(states, dtype=np.uint32)
    n = st.shape[0]
    lol = np.zeros(n); eue = np.zeros(n)
    lib().oracle_nonseq_states(len(cap), cap, len(load), load, n, st, lol, eue)
    return lol, eue


def nonseq_philox(cap, mttf, mttr, load, seed, i0, iters, want_states=True):
    cap = _d(cap); load = _d(load)
    thr = for_threshold(mttf, mttr)
    U = len(cap); W = (U + 31) // 32
    lol = np.zeros(iters); eue = np.zeros(iters)
    st = np.zeros((iters, W), dtype=np.uint32)
    lib().oracle_nonseq_philox(U, cap, thr, len(load), load, seed, i0, iters, lol, eue, st)
    return lol, eue, st


# ------------------------------------------------------------------------------ analytical
def copt_build(cap, for_rate, step):
    cap = _d(cap); q = _d(for_rate)
    max_len = int(np.ceil(cap.sum() / step)) + 2 * len(cap) + 8
    probs = np.zeros(max_len)
    n = lib().oracle_copt_build(len(cap), cap, q, float(step), probs, max_len)
    if n < 0:
        raise RuntimeError("COPT buffer too small")
    return probs[:n].copy()


def analytical(cap, for_rate, load, step=10.0):
    """run_analytical, PSA.jl:113-163 -> (lole, eue, probs)."""
    cap = _d(cap); load = _d(load)
    probs = copt_build(cap, for_rate, step)
    lole = C.c_double(); eue = C.c_double()
    total = 0.0
    for c in cap:                       # sum(g.capacity for g in gens), left to right
        total += float(c)
    lib().oracle_analytical_indices(probs, len(probs), float(step), total, len(load), load,
                                    C.byref(lole), C.byref(eue))
    return lole.value, eue.value, probs


def gaa_build(cap, for_rate, step):
    """generating_adequacy_assessment.jl:30-107 add_unit chain."""
    cap = _d(cap); q = _d(for_rate)
    probs = np.array([1.0])
    for c, qq in zip(cap, q):
        n2 = lib().oracle_copt_next_len(len(probs), float(c), float(step))
        new = np.zeros(n2)
        lib().oracle_gaa_add_unit(probs, len(probs), float(c), float(qq), float(step), new, n2)
        probs = new
    return probs


def gaa_indices(probs, step, ldc):
    ldc = _d(ldc)
    lole = C.c_double(); eue = C.c_double()
    lib().oracle_gaa_calculate_indices(_d(probs), len(probs), float(step), len(ldc), ldc,
                                       C.byref(lole), C.byref(eue))
    return lole.value, eue.value


def fd_build(cap, mtbf_h, mttr_h):
    """generating_adequacy_frequency.jl:23-34,53-149: cumulative P / F tables on a 1 MW grid."""
    P = np.array([1.0]); F = np.array([0.0])
    for c, mtbf, mttr in zip(cap, mtbf_h, mttr_h):
        lam = 8760.0 / mtbf
        mu = 8760.0 / mttr
        q = lam / (lam + mu)
        p = 1.0 - q
        cur_max = float(len(P) - 1)
        n_new = int(np.floor(cur_max + c)) + 1     # collect(0.0:1.0:new_max)
        Pn = np.zeros(n_new); Fn = np.zeros(n_new)
        lib().oracle_fd_add_unit(P, F, len(P), float(c), p, q, lam, Pn, Fn, n_new)
        P, F = Pn, Fn
    return P, F


def fd_evaluate(P, F, peak, installed):
    a = C.c_double(); b = C.c_double(); c = C.c_double()
    lib().oracle_fd_evaluate(_d(P), _d(F), len(P), float(peak), float(installed),
                             C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def markov2(mttf, mttr, dt, steps):
    out = np.zeros(steps)
    lib().oracle_markov2(1.0 / mttf, 1.0 / mttr, dt, steps, out)
    return out


def dtmc_capacity(mttf, mttr, cap, r):
    r = _d(r)
    T, U = r.shape
    out = np.zeros(T)
    lib().oracle_dtmc_capacity(U, _d(mttf), _d(mttr), _d(cap), T, r, out)
    return out


def load_factors(total_hours, weekly, daily, hourly):
    out = np.zeros(total_hours)
    lib().oracle_load_factors(total_hours, _d(weekly), _d(daily), _d(hourly).reshape(-1), out)
    return out


# ------------------------------------------------------------------------------- tail risk
def quantile_type7(x, alpha):
    """Julia Statistics.quantile default (type 7): position 1+(N-1)alpha, linear interpolation
    (SURVEY.md section 8 row a-12 build-side spec)."""
    xs = np.sort(np.asarray(x, dtype=np.float64))
    n = len(xs)
    pos = (n - 1) * alpha
    lo = int(np.floor(pos))
    hi = min(lo + 1, n - 1)
    g = pos - lo
    return xs[lo] + g * (xs[hi] - xs[lo])


def cvar(x, alpha):
    """Mean of the values >= VaR_alpha (a-12)."""
    x = np.asarray(x, dtype=np.float64)
    v = quantile_type7(x, alpha)
    tail = x[x >= v]
    return v, float(tail.mean())


# ------------------------------------------- hourly-resampled MC (tail_risk.jl:12-91), f-1
def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def detailed_mc_injected(cap, for_rate, maint_start, maint_weeks, energy_limit, base_load, lfu_std, unif, norm):
    """unif[n_years, H, U], norm[n_years, H] -> (yearly_lole[n_years], hourly_fail_count[H])."""
    unif = _d(unif); norm = _d(norm)
    n, H, U = unif.shape
    yl = np.zeros(n); hf = np.zeros(H)
    rc = lib().oracle_detailed_mc_injected(U, _d(cap), _d(for_rate), _i32(maint_start), _i32(maint_weeks),
                                           _d(energy_limit), H, _d(base_load), float(lfu_std), n, unif, norm, yl, hf)
    assert rc == 0
    return yl, hf


def detailed_mc_philox(cap, for_rate, maint_start, maint_weeks, energy_limit, base_load, lfu_std, seed, year0, n_years):
    q = _d(for_rate)
    thr = np.minimum(np.floor(q * 4294967296.0), 4294967295.0).astype(np.uint32)
    H = len(base_load)
    yl = np.zeros(n_years); hf = np.zeros(H)
    rc = lib().oracle_detailed_mc_philox(len(q), _d(cap), thr, _i32(maint_start), _i32(maint_weeks), _d(energy_limit),
                                         H, _d(base_load), float(lfu_std), seed, year0, n_years, yl, hf)
    assert rc == 0
    return yl, hf


def normal_u32x2(x1: int, x2: int) -> float:
    return float(lib().oracle_normal_u32x2(int(x1), int(x2)))


def schedule_maintenance(capacity, maintenance_weeks, weekly_peaks):
    """generating_adequacy_comprehensive.jl:86-112: units by capacity*weeks descending (stable), each placed
    at the start week that maximises the minimum reserve over its window (first maximum wins); returns the
    1-based start weeks (0 = no maintenance)."""
    cap = _d(capacity); mw = [int(w) for w in maintenance_weeks]; peaks = _d(weekly_peaks)
    avail = np.full(52, float(cap.sum()))
    order = sorted(range(len(cap)), key=lambda i: -(cap[i] * mw[i]))
    start = [0] * len(cap)
    for i in order:
        if mw[i] == 0:
            continue
        best, best_res = 1, -np.inf
        for s in range(1, 52 - mw[i] + 2):
            res = (avail[s - 1:s - 1 + mw[i]] - peaks[s - 1:s - 1 + mw[i]]).min()
            if res > best_res:
                best_res, best = res, s
        start[i] = best
        avail[best - 1:best - 1 + mw[i]] -= cap[i]
    return start


def expected_generation(probs, step, unit_cap, loads, lfu_sigma):
    """calculate_expected_generation, generating_adequacy_comprehensive.jl:118-142 (literal loops)."""
    p = _d(probs); ld = _d(loads)
    return float(lib().oracle_expected_generation(p, len(p), float(step), float(unit_cap), ld, len(ld), float(lfu_sigma)))


def lfu_hourly_risk(probs, step, loads, lfu_mw):
    """tail_risk.jl:124-136 hourly risk with the 7-step LFU table (literal loops)."""
    p = _d(probs); ld = _d(loads)
    out = np.zeros(len(ld))
    lib().oracle_lfu_hourly_risk(p, len(p), float(step), ld, len(ld), float(lfu_mw), out)
    return out
