"""B200-native HL1 generating-adequacy Monte Carlo (drop-in for the hot path of
Matrixeigs/PowerSystemsReliabilityAssessment, GeneratingAdequacy/PowerSystemAdequacy.jl).

The compute path is libpsra_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/psra_b200.h); this package is the thin host mirror of the reference's Julia API.
"""
from . import rts79  # noqa: F401
from ._lib import DISC_MATLAB, INIT_ALL_UP, INIT_STATIONARY, LIB_PATH  # noqa: F401
from .api import (DetailedGenerator, SystemParams, run_detailed_mc, run_monte_carlo, schedule_maintenance,  # noqa: F401
                  run_nonseq_until_beta, run_sequential_until_cov, run_detailed_analytical, update_elu,
                  calculate_expected_generation, get_lfu_distribution,
                  Engine, Generator, LoadModel, PsraError, ReliabilityResult, SequentialIndices,  # noqa: F401
                  compare_results, cumulative_series, export_results, export_results_mat, export_nonseq_results_mat, unit_importance, evaluate_risk, indices_from_raw, run_analytical,
                  run_non_sequential_mc, run_sequential_mc,
                  ISOLATED, INTERCONNECTED, AreaGenerator, TieLine, Area, System, run_fast_sequential_simulation)
