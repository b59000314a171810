"""Host-side mirror of the reference's exported API for the HL1 hot path.

Same names, argument meaning and return record as
GeneratingAdequacy/PowerSystemAdequacy.jl:8-10 (`Generator`, `LoadModel`,
`ReliabilityResult`, `run_analytical`, `run_non_sequential_mc`, `run_sequential_mc`,
`compare_results`); every engine body is one call into libpsra_b200.so (CUDA, sm_100a)
through the C ABI of include/psra_b200.h.  The Julia twin that a maintainer of the reference
would drop in is julia/PowerSystemAdequacyB200.jl; this Python module exists because Julia is
not installed in the build image, and it is what tests/ and bench.py drive.

No CPU fallback: constructing an `Engine` without the CUDA library / a GPU raises.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
import time
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import INIT_ALL_UP, INIT_STATIONARY


class PsraError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libpsra_b200 error {code}: {msg}")
        self.code = code


# ------------------------------------------------------------------ data model, PSA.jl:20-53
@dataclasses.dataclass
class Generator:
    """PSA.jl:20-37: lambda = 1/MTTF, mu = 1/MTTR, for_rate = lambda/(lambda+mu)."""
    id: int
    capacity: float
    mttf: float
    mttr: float
    lambda_: float = dataclasses.field(init=False)
    mu: float = dataclasses.field(init=False)
    for_rate: float = dataclasses.field(init=False)

    def __post_init__(self):
        self.lambda_ = 1.0 / self.mttf
        self.mu = 1.0 / self.mttr
        self.for_rate = self.lambda_ / (self.lambda_ + self.mu)


class LoadModel:
    """PSA.jl:39-45."""

    def __init__(self, hourly_load):
        self.hourly_load = np.ascontiguousarray(hourly_load, dtype=np.float64)
        self.peak_load = float(self.hourly_load.max())


@dataclasses.dataclass
class ReliabilityResult:
    """PSA.jl:47-53."""
    method: str
    lole_hours_yr: float
    eue_mwh_yr: float
    computation_time: float
    convergence_history: np.ndarray


@dataclasses.dataclass
class SequentialIndices:
    """Indices of a sequential run (Montecarlo_seq/seqMain.m:160-176,210-213 definitions)."""
    years: int
    lole: float            # h/yr   = mean(DLC)
    eens: float            # MWh/yr = mean(ENS)
    lolf: float            # occ/yr = mean(NLC), calnlc.m:22-34
    lold: float            # h/occ  = sum(DLC)/sum(NLC)
    lole_se: float         # standard errors of the two means
    eens_se: float
    p_loss_year: float     # share of years with any loss of load
    cov_eens: float        # seqMain.m:183-186 CoV = std(ENS)/(mean*sqrt(n))
    events: int
    kernel_ms: float
    redone: int = 0                            # chains the library replayed with the generic kernel (exact either way)
    lol_hours: Optional[np.ndarray] = None     # per-year vectors when requested
    ens: Optional[np.ndarray] = None
    entries: Optional[np.ndarray] = None
    fail_count: Optional[np.ndarray] = None
    group_lol: Optional[np.ndarray] = None
    history: Optional[np.ndarray] = None       # running mean of LOL hours every `history` years (device-computed)
    raw: Optional[dict] = None                 # exact integer accumulators (shardable)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _fixed_load(load_mw, scale: float, mode: str) -> np.ndarray:
    """Hourly load in MW -> the fixed-point grid.  The reference tests `cap_avail < load` in Float64 (PSA.jl:192,253) and
    the capacities are whole grid values here, so for a load L between two grid values the loss test is decided by
    ceil(L): c < L  <=>  c < ceil(L).  mode:
      "ceil"   (default) -- ceil to the grid: every loss-of-load hour of the Float64 comparison is kept exactly
                            (LOLE / LOLF / durations unbiased); the deficit load - cap is then over-stated by less than
                            one grid unit per loss hour (EENS bias < LOLE / fp_scale MWh per year) -- choose a finer
                            fp_scale to shrink it, or feed loads that are whole grid values (no change at all);
      "rint"             -- round to nearest: EENS unbiased to first order, but loss hours with a load within half a
                            grid unit above a capacity level are missed (RTS-79: LOLE 9.3677 instead of 9.3941 h/yr);
      "strict"           -- raise unless the load is representable exactly."""
    v = np.asarray(load_mw, dtype=np.float64) * scale
    r = np.rint(v)
    if mode == "strict":
        if not (np.abs(v - r) <= 1e-6).all():
            raise ValueError(f"load is not representable at fp_scale={scale}; choose a finer scale")
    elif mode == "ceil":
        # at fp_scale = 1 this is the Float64 comparison itself (a load one ulp above a whole number does lose load at
        # that capacity); a scale factor adds its own rounding error, forgiven up to a few ulps
        tol = 0.0 if scale == 1.0 else 8.0 * np.finfo(np.float64).eps * np.abs(v)
        r = np.where(np.abs(v - r) <= tol, r, np.ceil(v))
    elif mode != "rint":
        raise ValueError("load mode must be 'ceil', 'rint' or 'strict'")
    if r.size and (r.min() < 0 or r.max() > 0x3fffffff):
        raise ValueError(f"load out of the int32 fixed-point range at fp_scale={scale}")
    return np.ascontiguousarray(r, dtype=np.int32)


def _fixed(values, scale: float, what: str, strict: bool) -> np.ndarray:
    v = np.asarray(values, dtype=np.float64) * scale
    r = np.rint(v)
    if strict and not np.allclose(v, r, rtol=0, atol=1e-6):
        raise ValueError(f"{what} is not representable at fp_scale={scale}; choose a finer scale "
                         "or pass strict=False to round to the grid")
    if r.size and (r.min() < 0 or r.max() > 0x3fffffff):
        raise ValueError(f"{what} out of the int32 fixed-point range at fp_scale={scale}")
    return np.ascontiguousarray(r, dtype=np.int32)


class Engine:
    """One libpsra_b200 handle = one CUDA device + stream."""

    def __init__(self, device: int = 0, warps_per_block: int = 0, seg_hours: int = 0, blocks_per_sm: int = 0,
                 force_generic: bool = False, unpacked_words: bool = False, force_team: bool = False,
                 static_blocks: int = 0, ngpus: int = 1, ev_cap: int = 0, tail_bins: int = 0, single_stream: bool = False):
        """ngpus > 1: the handle spans the devices device .. device + ngpus - 1 of this process; the Monte Carlo calls
        shard their year / sample range over them and combine the integers with NCCL inside the library."""
        self._L = _lib.load()
        self._h = C.c_void_p()
        cfg = _lib.Config(device=device, warps_per_block=warps_per_block, seg_hours=seg_hours,
                          blocks_per_sm=blocks_per_sm, ngpus=int(ngpus), ev_cap=int(ev_cap), tail_bins=int(tail_bins))
        cfg.reserved[0] = 1 if force_generic else 0
        cfg.reserved[1] = int(unpacked_words)      # 1: seq_fast.cu keeps its word sums unpacked (cross-check variant)
        cfg.reserved[2] = 1 if force_team else 0
        cfg.reserved[3] = int(static_blocks)
        cfg.reserved2[0] = 1 if single_stream else 0   # chunked history runs: all launches on one stream (comparison)
        self.ngpus = max(1, int(ngpus))
        rc = self._L.psra_create(C.byref(self._h), C.byref(cfg))
        if rc != 0:
            msg = self._L.psra_last_error(self._h).decode() if self._h else "psra_create failed"
            if self._h:
                self._L.psra_destroy(self._h)
                self._h = C.c_void_p()
            raise PsraError(rc, msg)
        self.fp_scale = 1.0
        self.n_units = 0
        self.n_hours = 0

    # ---- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._L.psra_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise PsraError(rc, self._L.psra_last_error(self._h).decode())

    @property
    def stream(self) -> int:
        return int(self._L.psra_stream(self._h))

    def device_info(self):
        sm = C.c_int32(); khz = C.c_int32()
        self._check(self._L.psra_device_info(self._h, C.byref(sm), C.byref(khz)))
        return sm.value, khz.value

    def last_counters(self):
        """Diagnostic counters of the last Monte Carlo call (waves, Philox jobs, ...)."""
        out = (C.c_uint64 * 16)()
        self._check(self._L.psra_last_counters(self._h, out, 16))
        v = list(out)
        return dict(waves=v[9], jobs=v[10], ahead_jobs=v[11], resolved_runs=v[12], pend_max=v[13], events=v[7], redone=v[14])

    # ---- system data
    def set_system(self, capacity_mw, mttf_h, mttr_h, fp_scale: float = 1.0, strict: bool = True):
        cap = _fixed(capacity_mw, fp_scale, "capacity", strict)
        mttf = np.ascontiguousarray(mttf_h, dtype=np.float64)
        mttr = np.ascontiguousarray(mttr_h, dtype=np.float64)
        self._check(self._L.psra_set_system(self._h, _ptr(cap), _ptr(mttf), _ptr(mttr), len(cap)))
        self.fp_scale = float(fp_scale)
        self.n_units = len(cap)

    def set_load(self, load_mw, strict: bool = False, mode: Optional[str] = None):
        """mode: see _fixed_load ("ceil" keeps the Float64 loss test exact; default).  strict=True == mode="strict"."""
        load = _fixed_load(load_mw, self.fp_scale, mode or ("strict" if strict else "ceil"))
        self._check(self._L.psra_set_load(self._h, _ptr(load), len(load)))
        self.n_hours = len(load)
        return load

    def set_generators(self, gens: Sequence[Generator], load: LoadModel, fp_scale: float = 1.0,
                       strict: bool = True, load_mode: str = "ceil"):
        self.set_system([g.capacity for g in gens], [g.mttf for g in gens], [g.mttr for g in gens],
                        fp_scale, strict)
        return self.set_load(load.hourly_load, mode=load_mode)

    # ---- sequential MC
    def _seq_outputs(self, n, per_year, fail_count, group, keep, history=False, tail_hist=False):
        o = _lib.SeqOutputs()
        o.tail_hist = 1 if tail_hist else 0
        bufs = {}
        if history:
            bufs["history"] = np.empty(n // history, dtype=np.float64)     # fully overwritten by the library
            o.history = _ptr(bufs["history"]); o.group = history
        if per_year:
            bufs["lol_hours"] = np.zeros(n, dtype=np.uint32)
            bufs["ens"] = np.zeros(n, dtype=np.int64)
            bufs["entries"] = np.zeros(n, dtype=np.uint32)
            o.lol_hours = _ptr(bufs["lol_hours"]); o.ens_fp = _ptr(bufs["ens"]); o.entries = _ptr(bufs["entries"])
        if fail_count:
            bufs["fail_count"] = np.zeros(self.n_hours, dtype=np.uint32)
            o.fail_count = _ptr(bufs["fail_count"])
        if group:
            bufs["group_lol"] = np.zeros((n + group - 1) // group, dtype=np.int64)
            o.group_lol = _ptr(bufs["group_lol"]); o.group = group
        o.keep_on_device = 1 if keep else 0
        return o, bufs

    def _seq_result(self, s: _lib.SeqSummary, bufs) -> SequentialIndices:
        raw = dict(years=s.years, sum_lol_hours=s.sum_lol_hours, sum_ens_fp=s.sum_ens_fp,
                   sum_entries=s.sum_entries, years_with_loss=s.years_with_loss, sum_lol_sq=s.sum_lol_sq,
                   sum_ens_sq=(s.sum_ens_sq_hi << 64) | s.sum_ens_sq_lo, events=s.events)
        r = indices_from_raw(raw, self.fp_scale)
        r.kernel_ms = float(s.kernel_ms)
        r.redone = int(s.redone)
        sc = self.fp_scale
        r.lol_hours = bufs.get("lol_hours")
        r.ens = None if "ens" not in bufs else bufs["ens"] / sc
        r.entries = bufs.get("entries")
        r.fail_count = bufs.get("fail_count")
        r.group_lol = bufs.get("group_lol")
        r.history = bufs.get("history")
        if "ens" in bufs:
            raw["ens_fp_vector"] = bufs["ens"]
        return r

    def seq_mc(self, years: int, seed: int = 42, year0: int = 0, init_mode: int = INIT_STATIONARY,
               years_per_chain: int = 1, per_year: bool = False, fail_count: bool = False,
               group: int = 0, keep_on_device: bool = False, history: int = 0, tail_hist: bool = False) -> SequentialIndices:
        """tail_hist=True: the kernel also counts the years by their ENS (1 fixed-point MWh bins) on the device;
        Engine.tail() then gives exact VaR / CVaR without any per-year vector."""
        if history and group and history != group:
            raise ValueError("history and group must use the same cadence")
        o, bufs = self._seq_outputs(years, per_year, fail_count, group, keep_on_device, history, tail_hist)
        s = _lib.SeqSummary()
        self._check(self._L.psra_seq_mc(self._h, year0, years, seed, init_mode, years_per_chain,
                                        C.byref(o), C.byref(s)))
        return self._seq_result(s, bufs)

    def seq_unit_importance(self, years: int, seed: int = 42, year0: int = 0, init_mode: int = INIT_STATIONARY,
                            years_per_chain: int = 1, per_year: bool = False):
        """Montecarlo_seq/seqMain.m:140-150,225-231 at HL1 (weak-point detection): returns (importance, down_in_loss,
        indices) -- down_in_loss[u] = simulated hours with loss of load in which unit u is DOWN, importance =
        down_in_loss / total loss hours = the reference's comp_importance (P(unit down | system failure))."""
        o, bufs = self._seq_outputs(years, per_year, False, 0, False, 0)
        s = _lib.SeqSummary()
        cnt = np.zeros(self.n_units, dtype=np.uint64)
        self._check(self._L.psra_seq_unit_importance(self._h, year0, years, seed, init_mode, years_per_chain, _ptr(cnt),
                                                     C.byref(o), C.byref(s)))
        r = self._seq_result(s, bufs)
        tot = int(s.sum_lol_hours)
        imp = cnt.astype(np.float64) / tot if tot > 0 else np.zeros(self.n_units)
        return imp, cnt, r

    def seq_eval_injected(self, durations, years_per_chain: int = 1, fail_count: bool = False,
                          group: int = 0) -> SequentialIndices:
        """durations[nchains, U, K] (k = 0 initial TTF, then TTR, TTF, ...)."""
        d = np.ascontiguousarray(durations, dtype=np.float64)
        if d.ndim == 2:
            d = d[None]
        nchains, U, K = d.shape
        if U != self.n_units:
            raise ValueError("durations.shape[1] must equal the unit count")
        n = nchains * years_per_chain
        o, bufs = self._seq_outputs(n, True, fail_count, group, False)
        s = _lib.SeqSummary()
        self._check(self._L.psra_seq_eval_injected(self._h, _ptr(d), nchains, years_per_chain, K,
                                                   C.byref(o), C.byref(s)))
        return self._seq_result(s, bufs)

    # ---- non-sequential MC
    def _ns_outputs(self, n, per_sample, states, group, history=0):
        o = _lib.NonseqOutputs()
        bufs = {}
        if history:
            bufs["history"] = np.empty(n // history, dtype=np.float64)
            o.history = _ptr(bufs["history"]); o.group = history
        if per_sample:
            bufs["lol_hours"] = np.zeros(n, dtype=np.uint32)
            bufs["ens"] = np.zeros(n, dtype=np.int64)
            bufs["cap"] = np.zeros(n, dtype=np.int32)
            o.lol_hours = _ptr(bufs["lol_hours"]); o.ens_fp = _ptr(bufs["ens"]); o.cap_avail = _ptr(bufs["cap"])
        if states:
            bufs["states"] = np.zeros((n, (self.n_units + 31) // 32), dtype=np.uint32)
            o.states = _ptr(bufs["states"])
        if group:
            bufs["group_lol"] = np.zeros((n + group - 1) // group, dtype=np.int64)
            o.group_lol = _ptr(bufs["group_lol"]); o.group = group
        return o, bufs

    def _ns_result(self, s: _lib.NonseqSummary, bufs):
        n = max(s.samples, 1)
        sc = self.fp_scale
        e2 = (s.sum_ens_sq_hi << 64) | s.sum_ens_sq_lo
        mean_l = s.sum_lol_hours / n
        mean_e = s.sum_ens_fp / n
        var_l = max(s.sum_lol_sq / n - mean_l * mean_l, 0.0)
        var_e = max(e2 / n - mean_e * mean_e, 0.0)
        out = dict(samples=s.samples, lole=mean_l, eue=mean_e / sc, lole_se=math.sqrt(var_l / n),
                   eue_se=math.sqrt(var_e / n) / sc, p_loss=s.samples_with_loss / n,
                   kernel_ms=float(s.kernel_ms),
                   raw=dict(samples=s.samples, sum_lol_hours=s.sum_lol_hours, sum_ens_fp=s.sum_ens_fp,
                            samples_with_loss=s.samples_with_loss, sum_lol_sq=s.sum_lol_sq, sum_ens_sq=e2))
        out.update(bufs)
        return out

    def nonseq_mc(self, samples: int, seed: int = 42, sample0: int = 0, per_sample: bool = False,
                  states: bool = False, group: int = 0, history: int = 0):
        o, bufs = self._ns_outputs(samples, per_sample, states, group, history)
        s = _lib.NonseqSummary()
        self._check(self._L.psra_nonseq_mc(self._h, sample0, samples, seed, C.byref(o), C.byref(s)))
        return self._ns_result(s, bufs)

    def nonseq_eval_states(self, packed_states, group: int = 0):
        st = np.ascontiguousarray(packed_states, dtype=np.uint32)
        if st.ndim == 1:
            st = st[:, None]
        n = st.shape[0]
        o, bufs = self._ns_outputs(n, True, True, group)
        s = _lib.NonseqSummary()
        self._check(self._L.psra_nonseq_eval_states(self._h, _ptr(st), n, C.byref(o), C.byref(s)))
        return self._ns_result(s, bufs)

    def nonseq_eval_uniforms(self, r, group: int = 0):
        r = np.ascontiguousarray(r, dtype=np.float64)
        n, U = r.shape
        if U != self.n_units:
            raise ValueError("r.shape[1] must equal the unit count")
        o, bufs = self._ns_outputs(n, True, True, group)
        s = _lib.NonseqSummary()
        self._check(self._L.psra_nonseq_eval_uniforms(self._h, _ptr(r), n, C.byref(o), C.byref(s)))
        return self._ns_result(s, bufs)

    # ---- analytical
    def copt(self, capacity_mw, for_rate, step: float) -> np.ndarray:
        cap = np.ascontiguousarray(capacity_mw, dtype=np.float64)
        q = np.ascontiguousarray(for_rate, dtype=np.float64)
        max_len = int(math.ceil(cap.sum() / step)) + 2 * len(cap) + 8
        probs = np.zeros(max_len)
        n = C.c_int32()
        self._check(self._L.psra_copt(self._h, _ptr(cap), _ptr(q), len(cap), float(step), _ptr(probs),
                                      max_len, C.byref(n)))
        return probs[:n.value].copy()

    def copt_indices(self, probs, step: float, total_installed: float, load_mw):
        p = np.ascontiguousarray(probs, dtype=np.float64)
        ld = np.ascontiguousarray(load_mw, dtype=np.float64)
        a = C.c_double(); b = C.c_double()
        self._check(self._L.psra_copt_indices(self._h, _ptr(p), len(p), float(step), float(total_installed),
                                              _ptr(ld), len(ld), C.byref(a), C.byref(b)))
        return a.value, b.value

    def copt_indices_strict(self, probs, step: float, ldc_mw):
        p = np.ascontiguousarray(probs, dtype=np.float64)
        ld = np.ascontiguousarray(ldc_mw, dtype=np.float64)
        a = C.c_double(); b = C.c_double()
        self._check(self._L.psra_copt_indices_strict(self._h, _ptr(p), len(p), float(step), _ptr(ld), len(ld),
                                                     C.byref(a), C.byref(b)))
        return a.value, b.value

    def fd_recursion(self, capacity_mw, mtbf_h, mttr_h):
        cap = np.ascontiguousarray(capacity_mw, dtype=np.float64)
        a = np.ascontiguousarray(mtbf_h, dtype=np.float64)
        b = np.ascontiguousarray(mttr_h, dtype=np.float64)
        max_len = int(math.floor(cap.sum())) + len(cap) + 8
        P = np.zeros(max_len); F = np.zeros(max_len)
        n = C.c_int32()
        self._check(self._L.psra_fd_recursion(self._h, _ptr(cap), _ptr(a), _ptr(b), len(cap), _ptr(P), _ptr(F),
                                              max_len, C.byref(n)))
        return P[:n.value].copy(), F[:n.value].copy()

    def markov2(self, mttf: float, mttr: float, dt: float = 1.0, steps: int = 200) -> np.ndarray:
        out = np.zeros(steps)
        self._check(self._L.psra_markov2(self._h, 1.0 / mttf, 1.0 / mttr, float(dt), steps, _ptr(out)))
        return out

    def dtmc_capacity(self, mttf_h, mttr_h, capacity_mw, r) -> np.ndarray:
        r = np.ascontiguousarray(r, dtype=np.float64)
        T, U = r.shape
        a = np.ascontiguousarray(mttf_h, dtype=np.float64)
        b = np.ascontiguousarray(mttr_h, dtype=np.float64)
        c = np.ascontiguousarray(capacity_mw, dtype=np.float64)
        out = np.zeros(T)
        self._check(self._L.psra_dtmc_capacity(self._h, _ptr(a), _ptr(b), _ptr(c), U, _ptr(r), T, _ptr(out)))
        return out

    # ---- hourly-resampled MC with maintenance / LFU / ELU (tail_risk.jl:12-91)
    def _detailed_sys(self, gens):
        cap = np.ascontiguousarray([g.capacity for g in gens], dtype=np.float64)
        q = np.ascontiguousarray([g.for_rate for g in gens], dtype=np.float64)
        ms = np.ascontiguousarray([g.scheduled_outage_start for g in gens], dtype=np.int32)
        mw = np.ascontiguousarray([g.maintenance_weeks for g in gens], dtype=np.int32)
        el = np.ascontiguousarray([g.energy_limit for g in gens], dtype=np.float64)
        sys = _lib.DetailedSystem(_ptr(cap), _ptr(q), _ptr(ms), _ptr(mw), _ptr(el), len(gens), 0)
        return sys, (cap, q, ms, mw, el)

    def detailed_mc(self, gens, base_load, lfu_std: float, n_years: int, seed: int = 42, year0: int = 0):
        """psra_detailed_mc (tail_risk.jl:12-91); at most 30 units (the kernel keeps the per-unit energy state in registers)."""
        sys, keep = self._detailed_sys(gens)
        bl = np.ascontiguousarray(base_load, dtype=np.float64)
        yl = np.zeros(n_years, dtype=np.uint32); hf = np.zeros(len(bl), dtype=np.uint32)
        ms = C.c_float()
        self._check(self._L.psra_detailed_mc(self._h, C.byref(sys), _ptr(bl), len(bl), float(lfu_std), year0,
                                             n_years, seed, _ptr(yl), _ptr(hf), C.byref(ms)))
        return yl, hf, ms.value

    def detailed_eval_injected(self, gens, base_load, lfu_std: float, uniforms, normals):
        """Same loop on injected rand() / randn(): uniforms[year][hour][unit] (a value for EVERY unit; the reference draws
        none for a unit on maintenance, tail_risk.jl:39-44 -- lay a recorded stream out accordingly), normals[year][hour]."""
        sys, keep = self._detailed_sys(gens)
        bl = np.ascontiguousarray(base_load, dtype=np.float64)
        un = np.ascontiguousarray(uniforms, dtype=np.float64); no = np.ascontiguousarray(normals, dtype=np.float64)
        n = un.shape[0]
        yl = np.zeros(n, dtype=np.uint32); hf = np.zeros(len(bl), dtype=np.uint32)
        self._check(self._L.psra_detailed_eval_injected(self._h, C.byref(sys), _ptr(bl), len(bl), float(lfu_std), n,
                                                        _ptr(un), _ptr(no), _ptr(yl), _ptr(hf)))
        return yl, hf

    def failure_times(self, lam: float, n_samples: int = 10000, dt: float = 1.0, max_time: float = 5000.0, seed: int = 42,
                      uniforms=None) -> np.ndarray:
        """Markov_process.jl:39-60: the `failure_times` vector (components that outlive max_time are dropped, like the
        reference, which pushes nothing for them).  uniforms[n][K] injects rand()."""
        out = np.zeros(n_samples, dtype=np.float64)
        r = None; K = 0
        if uniforms is not None:
            r = np.ascontiguousarray(uniforms, dtype=np.float64)
            if r.ndim != 2 or r.shape[0] != n_samples:
                raise ValueError("uniforms must be [n_samples][K]")
            K = r.shape[1]
        self._check(self._L.psra_failure_times(self._h, float(lam), float(dt), float(max_time), n_samples, seed, _ptr(r), K, _ptr(out)))
        return out[out >= 0.0]

    # ---- multi-area adequacy (AdequacyAssessmentII.jl)
    def sampler_durations(self, mean_h: float, draws):
        """Sampler diagnostic: (ticks, e_bits) of 32-bit draws -- the duration in ticks of 2^-24 h that stands for
        -log(rand()) / rate (PowerSystemAdequacy.jl:224,243,246) and the bit pattern of the device's E = -ln(u)."""
        x = np.ascontiguousarray(draws, dtype=np.uint32)
        ticks = np.zeros(x.size, dtype=np.uint64)
        ebits = np.zeros(x.size, dtype=np.uint32)
        self._check(self._L.psra_sampler_durations(self._h, float(mean_h), _ptr(x), x.size, _ptr(ticks), _ptr(ebits)))
        return ticks, ebits

    def multi_area_mc(self, unit_area, cap, mttf, mttr, loads, topology, policy: int, years: int, seed: int = 42,
                      year0: int = 0, init_mode: int = INIT_STATIONARY, per_year: bool = False, fp_scale: float = 1.0,
                      strict: bool = True):
        """psra_multi_area_mc: loads[n_areas][H], topology[n_areas][n_areas] (System.topology_matrix).
        Leaves the engine's own system / load alone.  Returns dict(lole[A], eue[A], raw sums, optional per-year arrays)."""
        ua = np.ascontiguousarray(unit_area, dtype=np.int32)
        capi = _fixed(cap, fp_scale, "capacity", strict)
        mf = np.ascontiguousarray(mttf, dtype=np.float64); mr = np.ascontiguousarray(mttr, dtype=np.float64)
        ld = np.asarray(loads, dtype=np.float64)
        if ld.ndim != 2:
            raise ValueError("loads must be [n_areas][n_hours]")
        A, H = ld.shape
        ldi = _fixed_load(ld.reshape(-1), fp_scale, "ceil")        # like set_generators: the Float64 loss test stays exact
        topo = _fixed(np.asarray(topology, dtype=np.float64).reshape(-1), fp_scale, "tie capacity", strict)
        if len(topo) != A * A or not (len(ua) == len(capi) == len(mf) == len(mr)):
            raise ValueError("inconsistent multi-area system arrays")
        sys = _lib.AreaSystem(n_areas=A, n_units=len(capi), n_hours=H, reserved=0, unit_area=_ptr(ua), cap_fp=_ptr(capi),
                              mttf_h=_ptr(mf), mttr_h=_ptr(mr), load_fp=_ptr(ldi), topology_fp=_ptr(topo))
        o = _lib.AreaOutputs()
        lol = ens = None
        if per_year:
            lol = np.zeros((years, A), dtype=np.uint32); ens = np.zeros((years, A), dtype=np.int64)
            o.lol_hours = _ptr(lol); o.ens_fp = _ptr(ens)
        sm = _lib.AreaSummary()
        self._check(self._L.psra_multi_area_mc(self._h, C.byref(sys), int(policy), year0, years, seed, init_mode,
                                               C.byref(o), C.byref(sm)))
        # the library leaves the handle's single-area system / load untouched (the call uses private tables), so the
        # engine's unit count, hour count and scale still describe what set_system / set_load uploaded
        n = max(years, 1)
        sl = np.array(sm.sum_lol_hours[:A], dtype=np.int64); se = np.array(sm.sum_ens_fp[:A], dtype=np.int64)
        return dict(lole=sl / n, eue=se / n / fp_scale, sum_lol_hours=sl, sum_ens_fp=se, events=int(sm.events),
                    kernel_ms=float(sm.kernel_ms), lol_hours=lol, ens_fp=ens)

    # ---- tail risk
    def tail(self, values_fp=None, alphas=(0.95, 0.99), n_bins: int = 0, bin_width: int = 1):
        """VaR / CVaR (type-7 quantile, mean of values >= VaR) of integer per-year ENS.
        values_fp=None uses the vector kept on the device by seq_mc(keep_on_device=True)."""
        al = np.ascontiguousarray(alphas, dtype=np.float64)
        outs = (_lib.TailOut * len(al))()
        hist = np.zeros(max(n_bins, 1), dtype=np.int64)
        if values_fp is None:
            vp, n = None, 0
        else:
            v = np.ascontiguousarray(values_fp, dtype=np.int64)
            vp, n = _ptr(v), len(v)
        self._check(self._L.psra_tail(self._h, vp, n, _ptr(al), len(al), outs, _ptr(hist) if n_bins else None,
                                      n_bins, int(bin_width)))
        sc = self.fp_scale
        res = [dict(alpha=float(a), var=o.var / sc, cvar=o.cvar / sc, n_tail=o.n_tail, x_lo=o.x_lo, x_hi=o.x_hi)
               for a, o in zip(al, outs)]
        return (res, hist[:n_bins]) if n_bins else res


def _engine_tail_hist_export(self, max_bins: int = 1 << 24):
    """ENS histogram of the last seq_mc(tail_hist=True) as (counts[int64], meta[4] = years, years with loss, years beyond
    the range, their ENS sum): what ranks exchange (element-wise sum) for a cross-rank VaR / CVaR."""
    counts = np.zeros(max_bins, dtype=np.int64)
    meta = np.zeros(4, dtype=np.int64)
    n = C.c_int64()
    self._check(self._L.psra_tail_hist_export(self._h, _ptr(counts), max_bins, C.byref(n), _ptr(meta)))
    return counts[:n.value].copy(), meta


def _engine_tail_hist_import(self, counts, meta):
    c = np.ascontiguousarray(counts, dtype=np.int64)
    m = np.ascontiguousarray(meta, dtype=np.int64)
    self._check(self._L.psra_tail_hist_import(self._h, _ptr(c) if c.size else None, c.size, _ptr(m)))


Engine.tail_hist_export = _engine_tail_hist_export
Engine.tail_hist_import = _engine_tail_hist_import


def indices_from_raw(raw: dict, fp_scale: float = 1.0) -> SequentialIndices:
    """Indices from the exact integer accumulators; the accumulators of several shards
    (calls / GPUs) add component-wise, so this is also the post-allreduce step."""
    n = max(int(raw["years"]), 1)
    sl, se, sn = int(raw["sum_lol_hours"]), int(raw["sum_ens_fp"]), int(raw["sum_entries"])
    mean_l = sl / n
    mean_e = se / n
    var_l = max(int(raw["sum_lol_sq"]) / n - mean_l * mean_l, 0.0) * (n / max(n - 1, 1))
    var_e = max(int(raw["sum_ens_sq"]) / n - mean_e * mean_e, 0.0) * (n / max(n - 1, 1))
    cov = math.sqrt(var_e) / (mean_e * math.sqrt(n)) if mean_e > 0 else float("inf")
    return SequentialIndices(
        years=int(raw["years"]), lole=mean_l, eens=mean_e / fp_scale, lolf=sn / n,
        lold=(sl / sn) if sn else 0.0, lole_se=math.sqrt(var_l / n), eens_se=math.sqrt(var_e / n) / fp_scale,
        p_loss_year=int(raw["years_with_loss"]) / n, cov_eens=cov, events=int(raw.get("events", 0)),
        kernel_ms=0.0, raw=raw)


# ----------------------------------------------------------------- the reference entry points
_default_engine: Optional[Engine] = None


def default_engine() -> Engine:
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine()
    return _default_engine


def run_analytical(gens: Sequence[Generator], load: LoadModel, step_size: float = 10.0,
                   engine: Optional[Engine] = None) -> ReliabilityResult:
    """PSA.jl:113-163."""
    eng = engine or default_engine()
    t0 = time.time()
    cap = np.array([g.capacity for g in gens], dtype=np.float64)
    q = np.array([g.for_rate for g in gens], dtype=np.float64)
    probs = eng.copt(cap, q, step_size)
    total = 0.0
    for c in cap:            # sum(g.capacity for g in gens), left to right (PSA.jl:124)
        total += float(c)
    lole, eue = eng.copt_indices(probs, step_size, total, load.hourly_load)
    return ReliabilityResult("Analytical", lole, eue, time.time() - t0, np.zeros(0))


def run_non_sequential_mc(gens: Sequence[Generator], load: LoadModel, iterations: int, seed: int = 42,
                          fp_scale: float = 1.0, engine: Optional[Engine] = None) -> ReliabilityResult:
    """PSA.jl:169-208: history = running mean of LOLE every 100 iterations (:202-204)."""
    eng = engine or default_engine()
    t0 = time.time()
    eng.set_generators(gens, load, fp_scale)
    r = eng.nonseq_mc(iterations, seed=seed, history=100)
    return ReliabilityResult("Non-Sequential MC", r["lole"], r["eue"], time.time() - t0, r["history"])


def run_sequential_mc(gens: Sequence[Generator], load: LoadModel, years: int, seed: int = 42,
                      fp_scale: float = 1.0, init_mode: int = INIT_STATIONARY, years_per_chain: int = 1,
                      year0: int = 0, engine: Optional[Engine] = None, details: bool = False):
    """PSA.jl:214-269: history = running mean of LOLE every 10 years (:263-265).
    Default semantics differ from the reference in one documented way: the reference runs
    ONE chain that starts all-up and carries the unit states across all years (PSA.jl:223-224); the default here is
    independent years from the stationary law (init_mode=INIT_STATIONARY, years_per_chain=1) -- the same expectation
    without the all-up start bias, and what lets years shard over GPUs.  init_mode=INIT_ALL_UP, years_per_chain=years
    is the reference's chain.  An engine created with ngpus=G shards the years over G devices inside the library.
    year0 selects the shard [year0, year0+years) of the experiment `seed` (multi-process sharding / resume);
    details=True additionally returns the SequentialIndices (LOLF, LOLD, CIs, raw accumulators)."""
    eng = engine or default_engine()
    t0 = time.time()
    eng.set_generators(gens, load, fp_scale)
    r = eng.seq_mc(years, seed=seed, year0=year0, init_mode=init_mode, years_per_chain=years_per_chain, history=10)
    res = ReliabilityResult("Sequential MC", r.lole, r.eens, time.time() - t0, r.history)
    return (res, r) if details else res


def unit_importance(gens: Sequence[Generator], load: LoadModel, years: int, seed: int = 42, fp_scale: float = 1.0,
                    init_mode: int = INIT_STATIONARY, years_per_chain: int = 1, engine: Optional[Engine] = None):
    """Montecarlo_seq/seqMain.m:140-150,225-231 at HL1 (twin of julia unit_importance): (comp_importance,
    down_in_loss, SequentialIndices) for the generators of `gens`."""
    eng = engine or default_engine()
    eng.set_generators(gens, load, fp_scale)
    return eng.seq_unit_importance(years, seed=seed, init_mode=init_mode, years_per_chain=years_per_chain)


def compare_results(results: List[ReliabilityResult]) -> str:
    """PSA.jl:275-285 table (the Plots.jl figure of :288-297 is presentation, not reproduced)."""
    lines = ["", "==========================================", "       METHOD COMPARISON SUMMARY",
             "==========================================",
             "%-20s | %-10s | %-10s | %-10s" % ("Method", "LOLE(h/yr)", "EUE(MWh)", "Time(s)"), "-" * 60]
    for r in results:
        lines.append("%-20s | %-10.4f | %-10.2f | %-10.4f" % (r.method, r.lole_hours_yr, r.eue_mwh_yr,
                                                               r.computation_time))
    lines.append("-" * 60)
    text = "\n".join(lines)
    print(text)
    return text


# ---- multi-area adequacy: module AdequacyAssessmentFast (GeneratingAdequacy/AdequacyAssessmentII.jl)
ISOLATED = 0          # @enum SupportPolicy ISOLATED INTERCONNECTED (:63)
INTERCONNECTED = 1


@dataclasses.dataclass
class AreaGenerator:
    """AdequacyAssessmentII.jl:15-26 (the mutable state fields live on the device)."""
    id: str
    capacity: float
    mttf: float
    mttr: float


@dataclasses.dataclass
class TieLine:
    """AdequacyAssessmentII.jl:28-32; areas are 1-based like the reference."""
    from_area: int
    to_area: int
    capacity: float


@dataclasses.dataclass
class Area:
    """AdequacyAssessmentII.jl:34-39."""
    id: int
    name: str
    generators: List[AreaGenerator]
    hourly_load: np.ndarray


class System:
    """AdequacyAssessmentII.jl:41-61: topology_matrix[i, j] accumulates every tie line in both directions."""

    def __init__(self, areas: Sequence[Area], tie_lines: Sequence[TieLine]):
        self.areas = list(areas)
        self.tie_lines = list(tie_lines)
        n = len(self.areas)
        self.topology_matrix = np.zeros((n, n), dtype=np.float64)
        for ln in self.tie_lines:
            self.topology_matrix[ln.from_area - 1, ln.to_area - 1] += ln.capacity
            self.topology_matrix[ln.to_area - 1, ln.from_area - 1] += ln.capacity


def run_fast_sequential_simulation(sys: System, policy: int, n_years: int, seed: int = 42, fp_scale: float = 1.0,
                                   init_mode: int = INIT_STATIONARY, engine: Optional[Engine] = None, verbose: bool = True):
    """AdequacyAssessmentII.jl:185-250: returns [dict(area=name, lole=h/yr, eue=MWh/yr), ...] in area order.
    Years are independent (own Philox streams), see psra_multi_area_mc."""
    eng = engine or default_engine()
    t0 = time.time()
    if verbose:
        print("--- Running FAST Adequacy Assessment ---")
        print(f"Policy: {'ISOLATED' if policy == ISOLATED else 'INTERCONNECTED'} | Years: {n_years}")
    H = len(sys.areas[0].hourly_load)
    if any(len(a.hourly_load) != H for a in sys.areas):
        raise ValueError("all areas need load curves of the same length")
    ua = [i for i, a in enumerate(sys.areas) for _ in a.generators]
    gl = [g for a in sys.areas for g in a.generators]
    r = eng.multi_area_mc(ua, [g.capacity for g in gl], [g.mttf for g in gl], [g.mttr for g in gl],
                          np.stack([np.asarray(a.hourly_load, dtype=np.float64) for a in sys.areas]), sys.topology_matrix,
                          policy, n_years, seed=seed, init_mode=init_mode, fp_scale=fp_scale)
    if verbose:
        print(f"Simulation completed in {time.time() - t0:.2f} seconds.")
    return [dict(area=a.name, lole=float(r["lole"][i]), eue=float(r["eue"][i])) for i, a in enumerate(sys.areas)]


def evaluate_risk(cum_prob, cum_freq, peak_load: float, installed_cap: float):
    """generating_adequacy_frequency.jl:155-186 on the tables of Engine.fd_recursion (1 MW grid)."""
    reserve = installed_cap - peak_load
    for i in range(len(cum_prob)):
        if float(i) > reserve:
            lole_h = cum_prob[i] * 8760.0
            lolf = cum_freq[i]
            return lole_h, lolf, (lole_h / lolf if lolf > 0 else 0.0)
    return 0.0, 0.0, 0.0


# ------------------------------------------------ tail_risk.jl / MCvsMarkovProcess.jl entry points
@dataclasses.dataclass
class DetailedGenerator:
    """mutable struct Generator of generating_adequacy_comprehensive.jl:11-24 (the one tail_risk.jl includes)."""
    name: str
    capacity: float
    for_rate: float
    maintenance_weeks: int
    energy_limit: float = math.inf
    effective_q: float = None
    scheduled_outage_start: int = 0

    def __post_init__(self):
        if self.effective_q is None:
            self.effective_q = self.for_rate


def schedule_maintenance(gens: Sequence[DetailedGenerator], weekly_peaks) -> None:
    """schedule_maintenance!, generating_adequacy_comprehensive.jl:86-112 (greedy levelised reserve);
    sets g.scheduled_outage_start in place.  Host-side planning, O(52 * U)."""
    peaks = np.asarray(weekly_peaks, dtype=np.float64)
    total_installed = sum(g.capacity for g in gens)
    weekly_available = np.full(52, float(total_installed))
    for g in sorted(gens, key=lambda g: g.capacity * g.maintenance_weeks, reverse=True):
        if g.maintenance_weeks == 0:
            continue
        best_start, max_min_reserve = 1, -math.inf
        for start_w in range(1, 52 - g.maintenance_weeks + 2):
            w = slice(start_w - 1, start_w - 1 + g.maintenance_weeks)
            min_res = float((weekly_available[w] - peaks[w]).min())
            if min_res > max_min_reserve:
                max_min_reserve, best_start = min_res, start_w
        g.scheduled_outage_start = best_start
        weekly_available[best_start - 1:best_start - 1 + g.maintenance_weeks] -= g.capacity


def run_detailed_mc(gens: Sequence[DetailedGenerator], base_load, lfu_sigma_percent: float, n_years: int,
                    seed: int = 42, engine: Optional[Engine] = None):
    """tail_risk.jl:12-91 -> (yearly_lole_distribution, hourly_failure_prob)."""
    eng = engine or default_engine()
    base_load = np.asarray(base_load, dtype=np.float64)
    lfu_std_dev = float(base_load.max()) * (lfu_sigma_percent / 100.0)       # tail_risk.jl:22
    yl, hf, _ = eng.detailed_mc(gens, base_load, lfu_std_dev, n_years, seed=seed)
    return yl.astype(np.float64), hf.astype(np.float64) / n_years


@dataclasses.dataclass
class SystemParams:
    """MCvsMarkovProcess.jl:34-38."""
    step_size: float
    lfu_sigma_percent: float
    mc_years: int


def run_monte_carlo(gens: Sequence[DetailedGenerator], base_load, params: SystemParams, seed: int = 42,
                    engine: Optional[Engine] = None):
    """MCvsMarkovProcess.jl:210-284 -> (mean(yearly_lole), hourly_failures / years, yearly_lole)."""
    yl, prof = run_detailed_mc(gens, base_load, params.lfu_sigma_percent, params.mc_years, seed=seed, engine=engine)
    return float(yl.mean()), prof, yl


# ------------------------------------------------------------------ result series / export (SURVEY f-4)
def cumulative_series(r: SequentialIndices):
    """Montecarlo_seq/seqMain.m:180-186: results_cum.eens[i] = mean(ENS_1..i) and results_cum.cov[i] =
    std(ENS_1..i) / (eens[i] * sqrt(i)) (n-1 form; cov[0] = 0 like the reference's untouched first slot), from the
    per-year ENS vector of a run with per_year=True.  Host-side numpy over device-computed per-year integers."""
    if r.ens is None:
        raise ValueError("cumulative_series needs a run with per_year=True")
    x = np.asarray(r.ens, dtype=np.float64)
    n = np.arange(1, x.size + 1, dtype=np.float64)
    s1 = np.cumsum(x); s2 = np.cumsum(x * x)
    eens = s1 / n
    var = np.zeros_like(x)
    var[1:] = np.maximum(s2[1:] - s1[1:] * s1[1:] / n[1:], 0.0) / (n[1:] - 1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        cov = np.where(eens > 0, np.sqrt(var) / (eens * np.sqrt(n)), 0.0)
    cov[0] = 0.0
    return eens, cov


def export_results(prefix: str, r: SequentialIndices) -> List[str]:
    """Montecarlo_seq/seqMain.m:250-262 at HL1: `<prefix>_yearly.csv` with the per-year vectors the reference keeps
    in results_year (dlc = LOL hours, nlc = deficit entries, plc = dlc / H is left to the reader, ens in MWh) and the
    cumulative series of results_cum (eens, cov), plus `<prefix>_indices.csv` with the printed indices.  (The
    reference's nodal table and component ranking are HL2 outputs.)  Returns the written paths."""
    if r.lol_hours is None or r.ens is None or r.entries is None:
        raise ValueError("export_results needs a run with per_year=True")
    eens, cov = cumulative_series(r)
    yearly = prefix + "_yearly.csv"
    with open(yearly, "w") as f:
        f.write("year,dlc_hours,nlc_occ,ens_mwh,cum_eens_mwh_yr,cum_cov\n")
        for i in range(len(eens)):
            f.write(f"{i + 1},{int(r.lol_hours[i])},{int(r.entries[i])},{float(r.ens[i])!r},{float(eens[i])!r},{float(cov[i])!r}\n")
    idx = prefix + "_indices.csv"
    with open(idx, "w") as f:
        f.write("index,value\n")
        for k in ("years", "lole", "eens", "lolf", "lold", "lole_se", "eens_se", "p_loss_year", "cov_eens"):
            f.write(f"{k},{getattr(r, k)!r}\n")
    return [yearly, idx]


def export_results_mat(path: str, r: SequentialIndices, hours_per_year: int, comp_importance=None) -> str:
    """Montecarlo_seq/seqMain.m:260-262 at HL1: a MATLAB MAT-file (`save('seq_reliability_results.mat', 'results_year',
    'results_cum', 'nodal_eens_avg', 'comp_importance')`) with the same variable and field names --
    results_year.{plc, nlc, dlc, ens, dns} as 1 x years row vectors (seqMain.m:160-176: plc = dlc / HOURS_PER_YEAR, dns = ens /
    HOURS_PER_YEAR), results_cum.{eens, cov} (seqMain.m:180-186) and, when a weak-point run supplies it, comp_importance as a
    column (seqMain.m:225-231, generators only).  nodal_eens_avg is an HL2 output and is not written.  Needs scipy (host-side
    file format only); returns the path."""
    from scipy.io import savemat
    if r.lol_hours is None or r.ens is None or r.entries is None:
        raise ValueError("export_results_mat needs a run with per_year=True")
    H = float(hours_per_year)
    eens, cov = cumulative_series(r)
    row = lambda v: np.asarray(v, dtype=np.float64).reshape(1, -1)
    out = {"results_year": {"plc": row(r.lol_hours) / H, "nlc": row(r.entries), "dlc": row(r.lol_hours), "ens": row(r.ens),
                            "dns": row(r.ens) / H},
           "results_cum": {"eens": row(eens), "cov": row(cov)}}
    if comp_importance is not None:
        out["comp_importance"] = np.asarray(comp_importance, dtype=np.float64).reshape(-1, 1)
    savemat(path, out, oned_as="row")
    return path


def export_nonseq_results_mat(path: str, lole_history, edns_history, beta_history, comp_importance=None) -> str:
    """Montecarlo_nsq_single/nsqMain.m:403-405 at HL1: `save('reliability_results.mat', 'accumulated_edns',
    'accumulated_lole', 'nodal_eens', 'comp_importance', 'beta_history', 'edns_history')` -- the running series of a
    non-sequential run (one entry per batch: run_nonseq_until_beta returns the beta series, the caller accumulates LOLE / EDNS
    the same way) under the reference's variable names; nodal_eens is an HL2 output and is not written."""
    from scipy.io import savemat
    row = lambda v: np.asarray(v, dtype=np.float64).reshape(1, -1)
    out = {"accumulated_lole": row(lole_history), "accumulated_edns": row(edns_history), "edns_history": row(edns_history),
           "beta_history": row(beta_history)}
    if comp_importance is not None:
        out["comp_importance"] = np.asarray(comp_importance, dtype=np.float64).reshape(-1, 1)
    savemat(path, out, oned_as="row")
    return path


# ------------------------------------------------------------------ adaptive stopping (SURVEY f-2)
def run_sequential_until_cov(engine: Engine, cov_threshold: float = 0.05, batch_years: int = 1000,
                             max_years: int = 10_000_000, seed: int = 42, init_mode: int = INIT_STATIONARY,
                             years_per_chain: int = 1) -> SequentialIndices:
    """Montecarlo_seq/seqMain.m:183-197 stop rule: simulate until CoV = std(ENS_1..n)/(mean*sqrt(n)) drops below
    the threshold (and is > 0).  The reference tests the rule after every year; here after every batch of
    years (one kernel launch each), so the stop year is rounded up to a batch boundary."""
    tot = None
    done = 0
    while done < max_years:
        n = min(batch_years, max_years - done)
        r = engine.seq_mc(n, seed=seed, year0=done, init_mode=init_mode, years_per_chain=years_per_chain)
        tot = dict(r.raw) if tot is None else {k: tot[k] + r.raw[k] for k in tot}
        done += n
        idx = indices_from_raw(tot, engine.fp_scale)
        if 0 < idx.cov_eens < cov_threshold:
            break
    return indices_from_raw(tot, engine.fp_scale)


def run_nonseq_until_beta(engine: Engine, beta_threshold: float = 0.0017, batch: int = 100, max_samples: int = 100_000,
                          seed: int = 42):
    """Montecarlo_nsq_single/nsqMain.m:60-62,285-301 stop rule for the state sampler: beta =
    sqrt(sum (dns - EDNS)^2) / N / EDNS on the samples so far (peak-load mode: set_load([peak]) so that the
    per-sample ENS is the DNS); stops at beta < threshold or max_samples.  Returns (result dict, beta history)."""
    tot = None
    done = 0
    history = []
    while done < max_samples:
        n = min(batch, max_samples - done)
        r = engine.nonseq_mc(n, seed=seed, sample0=done)
        tot = dict(r["raw"]) if tot is None else {k: tot[k] + r["raw"][k] for k in tot}
        done += n
        N = tot["samples"]
        mean = tot["sum_ens_fp"] / N
        ss = tot["sum_ens_sq"] - N * mean * mean            # sum (dns - EDNS)^2
        beta = math.sqrt(max(ss, 0.0)) / N / mean if mean > 0 else math.inf
        history.append(beta)
        if beta < beta_threshold:
            break
    N = tot["samples"]
    out = dict(samples=N, plc=tot["samples_with_loss"] / N, lole=tot["sum_lol_hours"] / N,
               edns=tot["sum_ens_fp"] / N / engine.fp_scale, beta=history[-1], raw=tot)
    return out, np.array(history)


# ------------------------------ analytical side of tail_risk.jl (host-side planning; COPT tables on the GPU)
LFU_DISTRIBUTION = ((-3.0, 0.006), (-2.0, 0.061), (-1.0, 0.242), (0.0, 0.382), (1.0, 0.242), (2.0, 0.061), (3.0, 0.006))


def get_lfu_distribution():
    """7-step normal approximation, generating_adequacy_comprehensive.jl:76-80."""
    return list(LFU_DISTRIBUTION)


def _tail_weights(probs, step, thresholds):
    """For each threshold t: (sum of p_i with outage_i > t, sum of outage_i * p_i over the same states)."""
    n = len(probs)
    outages = np.arange(n) * step
    sp = np.concatenate([np.cumsum(probs[::-1])[::-1], [0.0]])
    sxp = np.concatenate([np.cumsum((outages * probs)[::-1])[::-1], [0.0]])
    first = np.searchsorted(outages, thresholds, side="right")        # first state with outage > t
    return sp[first], sxp[first], first


def calculate_expected_generation(probs, step: float, unit_cap: float, loads, lfu_sigma: float) -> float:
    """generating_adequacy_comprehensive.jl:118-142: expected energy an energy-limited unit of `unit_cap` MW must
    deliver on top of the rest of the system (COPT `probs` on the grid i*step), with the 7-step LFU.
    sum_i min(cap, outage_i - t) p_i over outage_i > t  =  [E(t) - E(t + cap)] with E(t) = sum (outage_i - t)+ p_i."""
    probs = np.asarray(probs, dtype=np.float64)
    loads = np.asarray(loads, dtype=np.float64)
    cap_rest = (len(probs) - 1) * step
    total = 0.0
    for z, pz in LFU_DISTRIBUTION:
        t = cap_rest - (loads + z * lfu_sigma)
        sp1, sxp1, _ = _tail_weights(probs, step, t)
        sp2, sxp2, _ = _tail_weights(probs, step, t + unit_cap)
        e1 = sxp1 - t * sp1                      # sum (outage - t)+ p
        e2 = sxp2 - (t + unit_cap) * sp2         # sum (outage - t - cap)+ p
        total += pz * float((e1 - e2).sum())
    return total


def update_elu(gens: Sequence[DetailedGenerator], loads, step_size: float, lfu_sigma: float,
               engine: Optional[Engine] = None) -> bool:
    """update_elu!, generating_adequacy_comprehensive.jl:144-175: effective FOR of the energy-limited units."""
    eng = engine or default_engine()
    changed = False
    for g in gens:
        if g.energy_limit == math.inf:
            continue
        rest = [og for og in gens if og is not g]
        probs = eng.copt([og.capacity for og in rest], [og.effective_q for og in rest], step_size)
        req_energy = calculate_expected_generation(probs, step_size, g.capacity, loads, lfu_sigma)
        new_q = g.for_rate
        if req_energy > g.energy_limit:
            deficit = req_energy - g.energy_limit
            new_q += deficit / (g.capacity * len(loads))
        new_q = min(new_q, 1.0)
        if abs(new_q - g.effective_q) > 1e-5:
            g.effective_q = new_q
            changed = True
    return changed


def run_detailed_analytical(gens: Sequence[DetailedGenerator], base_load, lfu_sigma_percent: float,
                            engine: Optional[Engine] = None):
    """tail_risk.jl:96-141 -> (sum(hourly_risk_profile), hourly_risk_profile): 5 ELU updates, 52 weekly COPTs
    (20 MW grid, units on maintenance left out), hourly risk with the 7-step LFU."""
    eng = engine or default_engine()
    base_load = np.asarray(base_load, dtype=np.float64)
    profile = np.zeros(len(base_load))
    step_size = 20.0
    lfu_mw = float(base_load.max()) * (lfu_sigma_percent / 100.0)
    for _ in range(5):
        update_elu(gens, base_load, step_size, lfu_mw, eng)
    for w in range(1, 53):
        week = [g for g in gens if not (w >= g.scheduled_outage_start and w < g.scheduled_outage_start + g.maintenance_weeks)]
        probs = eng.copt([g.capacity for g in week], [g.effective_q for g in week], step_size)
        installed = (len(probs) - 1) * step_size
        h0, h1 = (w - 1) * 168, min(w * 168, 8760)
        load = base_load[h0:h1]
        risk = np.zeros(len(load))
        for z, pz in LFU_DISTRIBUTION:
            sp, _, _ = _tail_weights(probs, step_size, installed - (load + z * lfu_mw))
            risk += sp * pz
        profile[h0:h1] = risk
    return float(profile.sum()), profile
