"""Tail-risk reduction (VaR / CVaR, histogram) on the GPU vs the numpy restatement of the a-12 spec."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 7, 1000, 100_003])
def test_quantiles_exact(engine, n):
    rng = np.random.default_rng(n)
    x = (rng.exponential(3000.0, n) * (rng.random(n) < 0.55)).astype(np.int64)
    if n > 5:
        x[:3] = [2**40 + 5, 0, 123456789012]
    engine.fp_scale = 1.0
    res = engine.tail(x, alphas=(0.0, 0.5, 0.95, 0.99, 1.0))
    for r in res:
        v, c = O.cvar(x, r["alpha"])
        assert r["var"] == v
        assert abs(r["cvar"] - c) <= 1e-12 * max(abs(c), 1.0)
        assert r["n_tail"] == int((x >= v).sum())


def test_tail_from_device_resident_run(engine, rts):
    """BASELINE config 4 shape at test size: per-year ENS stays on the device between the MC and
    the reduction; histogram as in tail_risk.jl:168."""
    engine.set_system(rts["cap"], rts["mttf"], rts["mttr"]); engine.set_load(rts["load_int"])
    r = engine.seq_mc(50_000, seed=11, per_year=True, keep_on_device=True)
    res, hist = engine.tail(None, alphas=(0.95, 0.99), n_bins=50, bin_width=1000)
    x = r.raw["ens_fp_vector"]
    for t in res:
        v, c = O.cvar(x, t["alpha"])
        assert t["var"] == v and abs(t["cvar"] - c) <= 1e-12 * c
    ref = np.bincount(np.minimum(x // 1000, 49), minlength=50)
    assert np.array_equal(hist, ref) and hist.sum() == 50_000
    assert 5000 < res[0]["var"] < 7500 and 12000 < res[1]["var"] < 16000    # BASELINE ballparks 6.3 / 13.8 GWh
