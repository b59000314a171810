"""IEEE RTS-79 HL1 data: unit table, hourly load curve, synthetic replicated system.

Sources (path:line under /root/reference):
  * MTTF / MTTR per generator row: Montecarlo_seq/case24_failrate.m:23-43 (33 rows, row 15 is
    the synchronous condenser, forced UP / ignored in Montecarlo_nsq_single/mc_sampling.m:40-41).
  * Load model: Montecarlo_seq/case24_loadprofile.m:18-73 (peak 2850 MW, weekly / daily /
    hourly factors) combined per Montecarlo_seq/anloducurve.m:24-88.
  * Unit capacities are NOT shipped by the reference (they live in MATPOWER's
    case24_ieee_rts, loaded at Montecarlo_seq/seqMain.m:32): the published IEEE RTS-79
    generator table in MATPOWER row order is embedded here (SURVEY.md section 8c [external]).

Pure host-side numpy; no CUDA, no oracle.
"""
from __future__ import annotations

import math

import numpy as np

HOURS_PER_YEAR = 8736  # 52 * 168, Montecarlo_seq/seqMain.m:38
PEAK_MW = 2850.0       # case24_loadprofile.m:18

# case24_failrate.m:23-43, 33 generator rows (row index 14, 0-based, = sync condenser)
_GEN_MTTF_33 = [
    450, 450, 1960, 1960, 450,
    450, 1960, 1960, 1200, 1200,
    1200, 950, 950, 950, 10000,
    2940, 2940, 2940, 2940, 2940,
    960, 960, 1100, 1100, 1980,
    1980, 1980, 1980, 1980, 1980,
    960, 960, 1150,
]
_GEN_MTTR_33 = [
    50, 50, 40, 40, 50,
    50, 40, 40, 50, 50,
    50, 50, 50, 50, 0.1,
    60, 60, 60, 60, 60,
    40, 40, 150, 150, 20,
    20, 20, 20, 20, 20,
    40, 40, 100,
]
# IEEE RTS-79 / MATPOWER case24_ieee_rts generator capacities, same row order [external]
_GEN_CAP_33 = [
    20, 20, 76, 76, 20,
    20, 76, 76, 100, 100,
    100, 197, 197, 197, 0,
    12, 12, 12, 12, 12,
    155, 155, 400, 400, 50,
    50, 50, 50, 50, 50,
    155, 155, 350,
]
SYNC_CONDENSER_ROW = 14  # 0-based (MATLAB index 15)

# case24_loadprofile.m:23-37
WEEKLY = np.array([
    0.862, 0.900, 0.878, 0.834, 0.880, 0.841, 0.832, 0.806,
    0.740, 0.737, 0.715, 0.727, 0.704, 0.750, 0.721, 0.800,
    0.754, 0.837, 0.870, 0.880, 0.856, 0.811, 0.900, 0.887,
    0.896, 0.861, 0.755, 0.816, 0.801, 0.880, 0.722, 0.776,
    0.800, 0.729, 0.726, 0.705, 0.780, 0.695, 0.724, 0.723,
    0.743, 0.744, 0.800, 0.881, 0.885, 0.909, 0.940, 0.890,
    0.942, 0.970, 1.000, 0.952,
])
# case24_loadprofile.m:41 (Mon..Sun)
DAILY = np.array([0.93, 1.00, 0.98, 0.96, 0.94, 0.77, 0.75])
# case24_loadprofile.m:48-73; columns: winter wkdy, winter wknd, summer wkdy, summer wknd,
# spring/fall wkdy, spring/fall wknd
HOURLY = np.array([
    [0.67, 0.78, 0.64, 0.74, 0.63, 0.75],
    [0.63, 0.72, 0.60, 0.70, 0.62, 0.73],
    [0.60, 0.68, 0.58, 0.66, 0.60, 0.69],
    [0.59, 0.66, 0.56, 0.65, 0.58, 0.66],
    [0.59, 0.64, 0.56, 0.64, 0.59, 0.65],
    [0.60, 0.65, 0.58, 0.62, 0.65, 0.65],
    [0.74, 0.66, 0.64, 0.62, 0.72, 0.68],
    [0.86, 0.70, 0.76, 0.66, 0.85, 0.74],
    [0.95, 0.80, 0.87, 0.81, 0.95, 0.83],
    [0.96, 0.88, 0.95, 0.86, 0.99, 0.89],
    [0.96, 0.90, 0.99, 0.91, 1.00, 0.92],
    [0.95, 0.91, 1.00, 0.93, 0.99, 0.94],
    [0.95, 0.90, 0.99, 0.93, 0.93, 0.91],
    [0.95, 0.88, 1.00, 0.92, 0.92, 0.90],
    [0.93, 0.87, 1.00, 0.91, 0.90, 0.90],
    [0.94, 0.87, 0.97, 0.91, 0.88, 0.86],
    [0.99, 0.91, 0.96, 0.92, 0.90, 0.85],
    [1.00, 1.00, 0.96, 0.94, 0.92, 0.88],
    [1.00, 0.99, 0.93, 0.95, 0.96, 0.92],
    [0.96, 0.97, 0.92, 0.95, 0.98, 1.00],
    [0.91, 0.94, 0.92, 1.00, 0.96, 0.97],
    [0.83, 0.92, 0.93, 0.93, 0.90, 0.95],
    [0.73, 0.87, 0.87, 0.88, 0.80, 0.90],
    [0.63, 0.81, 0.72, 0.80, 0.70, 0.85],
])


def units():
    """(capacity_mw, mttf_h, mttr_h) float64 arrays for the 32 generating units
    (sync condenser row dropped)."""
    keep = [i for i in range(33) if i != SYNC_CONDENSER_ROW]
    cap = np.array([_GEN_CAP_33[i] for i in keep], dtype=np.float64)
    mttf = np.array([_GEN_MTTF_33[i] for i in keep], dtype=np.float64)
    mttr = np.array([_GEN_MTTR_33[i] for i in keep], dtype=np.float64)
    return cap, mttf, mttr


def load_factors(total_hours: int = HOURS_PER_YEAR) -> np.ndarray:
    """Hourly factor week*day*hour, anloducurve.m:24-88 semantics (1-based hour h):
    week = ceil(h/168); winter weeks 1-8 and 44-52, summer 18-30; day = ceil(mod(h/24, 7)),
    0 -> 7; hour of day = mod(h, 24), 0 -> 24; column = 2*season + weekend."""
    if not 1 <= total_hours <= 52 * 168:
        raise ValueError("the weekly table covers 52 weeks (8736 h)")
    out = np.empty(total_hours, dtype=np.float64)
    for h in range(1, total_hours + 1):
        week = math.ceil(h / 168)
        if week <= 8 or week >= 44:
            season = 0
        elif 18 <= week <= 30:
            season = 1
        else:
            season = 2
        day = math.ceil(math.fmod(h / 24, 7))
        if day == 0:
            day = 7
        weekend = 0 if day <= 5 else 1
        hod = h % 24
        if hod == 0:
            hod = 24
        out[h - 1] = WEEKLY[week - 1] * DAILY[day - 1] * HOURLY[hod - 1, 2 * season + weekend]
    return out


def load_curve_mw(total_hours: int = HOURS_PER_YEAR, peak: float = PEAK_MW) -> np.ndarray:
    """Float64 MW curve: peak * factor (case24_loadprofile.m:18, anloducurve.m:87)."""
    return peak * load_factors(total_hours)


def load_curve_int(total_hours: int = HOURS_PER_YEAR, peak: float = PEAK_MW, scale: float = 1.0) -> np.ndarray:
    """int32 fixed-point curve rint(scale * MW) (round-half-even like Julia's round())."""
    return np.rint(scale * load_curve_mw(total_hours, peak)).astype(np.int32)


def synthetic_system(replicas: int = 32, load_scale: float = 37.0):
    """Config 5: RTS-79 unit table replicated `replicas` times (1024 units at 32) and the
    RTS curve scaled by `load_scale` then rounded to integer MW.  load_scale = 32 keeps the
    reserve margin (vacuous risk, LOLE ~ 1e-12 h/yr); 37 gives LOLE ~ 8 h/yr."""
    cap, mttf, mttr = units()
    cap = np.tile(cap, replicas)
    mttf = np.tile(mttf, replicas)
    mttr = np.tile(mttr, replicas)
    load = np.rint(load_scale * load_curve_mw()).astype(np.int32)
    return cap, mttf, mttr, load
