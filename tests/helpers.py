"""Shared input builders for the parity tests (seeded numpy, no reference access)."""
import numpy as np


def injected_durations(rng, mttf, mttr, nchains, K, scale=1.0):
    """durations[nchains, U, K]: k even -> time to failure ~ Exp(MTTF), k odd -> repair ~ Exp(MTTR)
    (the order PSA.jl:224,243,246 consumes them for a unit that starts UP)."""
    U = len(mttf)
    d = np.empty((nchains, U, K), dtype=np.float64)
    for k in range(K):
        mean = (mttf if k % 2 == 0 else mttr) * scale
        d[:, :, k] = rng.exponential(1.0, size=(nchains, U)) * mean[None, :]
    return np.maximum(d, 1e-12)


def draws_needed(mttf, mttr, hours, margin=2.0, floor=24):
    cyc = (np.asarray(mttf) + np.asarray(mttr)).min()
    return int(max(floor, margin * 2.0 * hours / cyc + floor))
