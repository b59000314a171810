"""ctypes binding of libpsra_b200.so (include/psra_b200.h).  No torch, no numpy math here:
plain pointers and sizes, exactly what a Julia `ccall` passes (julia/PowerSystemAdequacyB200.jl).

There is no CPU fallback: if the shared library is missing this module raises ImportError-like
RuntimeError on first use, and psra_create fails loudly without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpsra_b200.so")

PSRA_OK = 0
PSRA_E_INVALID = -1
PSRA_E_CUDA = -2
PSRA_E_OVERFLOW = -3
PSRA_E_NCCL = -4
INIT_ALL_UP = 0
INIT_STATIONARY = 1
DISC_MATLAB = 0x100

EXPORTS = [
    "psra_create", "psra_destroy", "psra_last_error", "psra_version", "psra_stream", "psra_device_info", "psra_last_counters",
    "psra_set_system", "psra_set_load", "psra_seq_mc", "psra_seq_eval_injected", "psra_nonseq_mc",
    "psra_nonseq_eval_states", "psra_nonseq_eval_uniforms", "psra_copt", "psra_copt_indices",
    "psra_copt_indices_strict", "psra_fd_recursion", "psra_markov2", "psra_dtmc_capacity", "psra_tail", "psra_detailed_mc", "psra_detailed_eval_injected",
    "psra_multi_area_mc", "psra_failure_times", "psra_sampler_durations", "psra_seq_unit_importance",
    "psra_tail_hist_export", "psra_tail_hist_import",
]


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("warps_per_block", C.c_int32), ("seg_hours", C.c_int32),
                ("blocks_per_sm", C.c_int32), ("reserved", C.c_int32 * 4), ("ngpus", C.c_int32), ("ev_cap", C.c_int32),
                ("tail_bins", C.c_int32), ("reserved2", C.c_int32 * 5)]


class SeqSummary(C.Structure):
    _fields_ = [("years", C.c_int64), ("sum_lol_hours", C.c_int64), ("sum_ens_fp", C.c_int64),
                ("sum_entries", C.c_int64), ("years_with_loss", C.c_int64), ("sum_lol_sq", C.c_uint64),
                ("sum_ens_sq_lo", C.c_uint64), ("sum_ens_sq_hi", C.c_uint64), ("events", C.c_uint64),
                ("kernel_ms", C.c_float), ("redone", C.c_int32)]


class SeqOutputs(C.Structure):
    _fields_ = [("lol_hours", C.c_void_p), ("ens_fp", C.c_void_p), ("entries", C.c_void_p),
                ("fail_count", C.c_void_p), ("group_lol", C.c_void_p), ("group", C.c_int32),
                ("keep_on_device", C.c_int32), ("history", C.c_void_p), ("tail_hist", C.c_int32),
                ("reserved", C.c_int32)]


class NonseqSummary(C.Structure):
    _fields_ = [("samples", C.c_int64), ("sum_lol_hours", C.c_int64), ("sum_ens_fp", C.c_int64),
                ("samples_with_loss", C.c_int64), ("sum_lol_sq", C.c_uint64), ("sum_ens_sq_lo", C.c_uint64),
                ("sum_ens_sq_hi", C.c_uint64), ("kernel_ms", C.c_float), ("reserved", C.c_int32)]


class NonseqOutputs(C.Structure):
    _fields_ = [("lol_hours", C.c_void_p), ("ens_fp", C.c_void_p), ("cap_avail", C.c_void_p),
                ("states", C.c_void_p), ("group_lol", C.c_void_p), ("group", C.c_int32),
                ("reserved", C.c_int32), ("history", C.c_void_p)]


class DetailedSystem(C.Structure):
    _fields_ = [("capacity_mw", C.c_void_p), ("for_rate", C.c_void_p), ("maint_start_week", C.c_void_p),
                ("maint_weeks", C.c_void_p), ("energy_limit_mwh", C.c_void_p), ("n_units", C.c_int32),
                ("reserved", C.c_int32)]


MAX_AREAS = 8


class AreaSystem(C.Structure):
    _fields_ = [("n_areas", C.c_int32), ("n_units", C.c_int32), ("n_hours", C.c_int32), ("reserved", C.c_int32),
                ("unit_area", C.c_void_p), ("cap_fp", C.c_void_p), ("mttf_h", C.c_void_p), ("mttr_h", C.c_void_p),
                ("load_fp", C.c_void_p), ("topology_fp", C.c_void_p)]


class AreaOutputs(C.Structure):
    _fields_ = [("lol_hours", C.c_void_p), ("ens_fp", C.c_void_p)]


class AreaSummary(C.Structure):
    _fields_ = [("years", C.c_int64), ("sum_lol_hours", C.c_int64 * MAX_AREAS), ("sum_ens_fp", C.c_int64 * MAX_AREAS),
                ("events", C.c_uint64), ("kernel_ms", C.c_float), ("n_areas", C.c_int32)]


class TailOut(C.Structure):
    _fields_ = [("var", C.c_double), ("cvar", C.c_double), ("n_tail", C.c_int64), ("x_lo", C.c_int64),
                ("x_hi", C.c_int64)]


_lib = None


def load():
    """Load the shared library and declare the prototypes.  Raises RuntimeError if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    L.psra_create.restype = C.c_int; L.psra_create.argtypes = [C.POINTER(vp), C.POINTER(Config)]
    L.psra_destroy.restype = None; L.psra_destroy.argtypes = [vp]
    L.psra_last_error.restype = C.c_char_p; L.psra_last_error.argtypes = [vp]
    L.psra_version.restype = C.c_int; L.psra_version.argtypes = []
    L.psra_stream.restype = u64; L.psra_stream.argtypes = [vp]
    L.psra_device_info.restype = C.c_int; L.psra_device_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    L.psra_last_counters.restype = C.c_int; L.psra_last_counters.argtypes = [vp, vp, i32]
    L.psra_set_system.restype = C.c_int; L.psra_set_system.argtypes = [vp, vp, vp, vp, i32]
    L.psra_set_load.restype = C.c_int; L.psra_set_load.argtypes = [vp, vp, i32]
    L.psra_seq_mc.restype = C.c_int
    L.psra_seq_mc.argtypes = [vp, i64, i64, u64, i32, i32, C.POINTER(SeqOutputs), C.POINTER(SeqSummary)]
    L.psra_seq_unit_importance.restype = C.c_int
    L.psra_seq_unit_importance.argtypes = [vp, i64, i64, u64, i32, i32, vp, C.POINTER(SeqOutputs), C.POINTER(SeqSummary)]
    L.psra_seq_eval_injected.restype = C.c_int
    L.psra_seq_eval_injected.argtypes = [vp, vp, i64, i32, i32, C.POINTER(SeqOutputs), C.POINTER(SeqSummary)]
    L.psra_nonseq_mc.restype = C.c_int
    L.psra_nonseq_mc.argtypes = [vp, i64, i64, u64, C.POINTER(NonseqOutputs), C.POINTER(NonseqSummary)]
    L.psra_nonseq_eval_states.restype = C.c_int
    L.psra_nonseq_eval_states.argtypes = [vp, vp, i64, C.POINTER(NonseqOutputs), C.POINTER(NonseqSummary)]
    L.psra_nonseq_eval_uniforms.restype = C.c_int
    L.psra_nonseq_eval_uniforms.argtypes = [vp, vp, i64, C.POINTER(NonseqOutputs), C.POINTER(NonseqSummary)]
    L.psra_copt.restype = C.c_int; L.psra_copt.argtypes = [vp, vp, vp, i32, dbl, vp, i32, C.POINTER(i32)]
    L.psra_copt_indices.restype = C.c_int
    L.psra_copt_indices.argtypes = [vp, vp, i32, dbl, dbl, vp, i32, C.POINTER(dbl), C.POINTER(dbl)]
    L.psra_copt_indices_strict.restype = C.c_int
    L.psra_copt_indices_strict.argtypes = [vp, vp, i32, dbl, vp, i32, C.POINTER(dbl), C.POINTER(dbl)]
    L.psra_fd_recursion.restype = C.c_int
    L.psra_fd_recursion.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, C.POINTER(i32)]
    L.psra_markov2.restype = C.c_int; L.psra_markov2.argtypes = [vp, dbl, dbl, dbl, i32, vp]
    L.psra_dtmc_capacity.restype = C.c_int; L.psra_dtmc_capacity.argtypes = [vp, vp, vp, vp, i32, vp, i32, vp]
    L.psra_failure_times.restype = C.c_int
    L.psra_failure_times.argtypes = [vp, dbl, dbl, dbl, i64, u64, vp, i32, vp]
    L.psra_sampler_durations.restype = C.c_int
    L.psra_sampler_durations.argtypes = [vp, C.c_float, vp, i64, vp, vp]
    L.psra_tail.restype = C.c_int
    L.psra_tail.argtypes = [vp, vp, i64, vp, i32, C.POINTER(TailOut), vp, i32, i64]
    L.psra_tail_hist_export.restype = C.c_int
    L.psra_tail_hist_export.argtypes = [vp, vp, i64, C.POINTER(i64), vp]
    L.psra_tail_hist_import.restype = C.c_int
    L.psra_tail_hist_import.argtypes = [vp, vp, i64, vp]
    L.psra_detailed_mc.restype = C.c_int
    L.psra_detailed_mc.argtypes = [vp, C.POINTER(DetailedSystem), vp, i32, dbl, i64, i64, u64, vp, vp, C.POINTER(C.c_float)]
    L.psra_multi_area_mc.restype = C.c_int
    L.psra_multi_area_mc.argtypes = [vp, C.POINTER(AreaSystem), i32, i64, i64, u64, i32, C.POINTER(AreaOutputs),
                                     C.POINTER(AreaSummary)]
    L.psra_detailed_eval_injected.restype = C.c_int
    L.psra_detailed_eval_injected.argtypes = [vp, C.POINTER(DetailedSystem), vp, i32, dbl, i64, vp, vp, vp, vp]
    _lib = L
    return L
