"""Golden vectors produced by the REFERENCE's own source text (needs /root/reference; run in the build container).

oracle/jl_transliterate.py cuts run_sequential_mc / run_non_sequential_mc / add_unit_convolution / run_analytical (and the
Generator constructor) out of GeneratingAdequacy/PowerSystemAdequacy.jl, transliterates them line by line into Python
and executes them on seeded inputs -- the three duration draws of run_sequential_mc (:224,243,246) read from per-unit
lists exactly as tools/patched_reference.jl does for a real Julia, rand() of run_non_sequential_mc replayed from a
recorded matrix.  Inputs and outputs go to tests/golden/ref_*.npz; tests/test_reference_pin.py holds the C oracle, the
hand transcription oracle/psa_literal.py and (on the GPU box, tests/test_gpu_golden.py) the CUDA path to them.
The reference text itself is not stored.

usage: python scripts/make_reference_golden.py [/root/reference]"""
import hashlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import jl_transliterate as J
from powersystemsreliabilityassessment_b200 import rts79

ref_root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = J.load_reference(ref_root)
sha = hashlib.sha256(src.encode("utf-8")).hexdigest()
out = os.path.join(ROOT, "tests", "golden")


def durations(rng, mttf, mttr, K, tiny_every=0):
    """D[u][k]: k even = time to failure (mean MTTF), k odd = time to repair -- the order :224,243,246 consume them."""
    U = len(mttf)
    D = np.empty((U, K))
    for u in range(U):
        for k in range(K):
            D[u, k] = -np.log(1.0 - rng.random()) * (mttf[u] if k % 2 == 0 else mttr[u])
    if tiny_every:
        D[:, tiny_every::tiny_every] *= 1e-3          # several transitions inside one hour (the `while ttf <= 0` loop)
    return D


cap, mttf, mttr = rts79.units()
cases = {}
# --- sequential, RTS-79, integer-MW load and the fractional MW curve
rng = np.random.default_rng(20261017)
for name, load, years in (("rts79_int", rts79.load_curve_int().astype(np.float64), 12), ("rts79_mw", rts79.load_curve_mw(), 10)):
    D = durations(rng, mttf, mttr, 800, tiny_every=7)
    res, lole, eue, hit = J.reference_sequential(src, cap, mttf, mttr, load, years, D)
    assert hit == [224, 243, 246], hit
    cases["seq_" + name] = dict(cap=cap, mttf=mttf, mttr=mttr, load=load, dur=D, lole=np.array(lole), eue=np.array(eue),
                                history=np.array(res.convergence_history), lole_hours_yr=res.lole_hours_yr, eue_mwh_yr=res.eue_mwh_yr)
    print(name, "LOLE", res.lole_hours_yr, "EUE", res.eue_mwh_yr, "history", res.convergence_history, flush=True)
# --- sequential, small random system (ragged hours, a unit larger than the load, frequent events)
U, H, years = 7, 500, 40
c7 = rng.integers(5, 60, U).astype(np.float64); f7 = rng.uniform(20.0, 200.0, U); r7 = rng.uniform(2.0, 30.0, U)
l7 = rng.integers(60, 170, H).astype(np.float64)
D = durations(rng, f7, r7, 2400, tiny_every=5)
res, lole, eue, hit = J.reference_sequential(src, c7, f7, r7, l7, years, D)
cases["seq_small"] = dict(cap=c7, mttf=f7, mttr=r7, load=l7, dur=D, lole=np.array(lole), eue=np.array(eue),
                          history=np.array(res.convergence_history), lole_hours_yr=res.lole_hours_yr, eue_mwh_yr=res.eue_mwh_yr)
print("small", res.lole_hours_yr, res.eue_mwh_yr, flush=True)
# --- non-sequential, RTS-79, 300 iterations (three history entries)
r = rng.random((300, len(cap)))
load = rts79.load_curve_mw()
res, lole, eue, q = J.reference_non_sequential(src, cap, mttf, mttr, load, 300, r.reshape(-1))
cases["nonseq_rts79_mw"] = dict(cap=cap, mttf=mttf, mttr=mttr, load=load, r=r, for_rate=np.array(q), lole=np.array(lole), eue=np.array(eue),
                                history=np.array(res.convergence_history), lole_hours_yr=res.lole_hours_yr, eue_mwh_yr=res.eue_mwh_yr)
print("nonseq", res.lole_hours_yr, res.eue_mwh_yr, res.convergence_history, flush=True)
# --- analytical, RTS-79, step 10 (exact-match branch) and step 7 (interpolation branch)
for step in (10.0, 7.0):
    res, probs, q = J.reference_analytical(src, cap, mttf, mttr, rts79.load_curve_mw(), step)
    cases[f"analytical_step{int(step)}"] = dict(cap=cap, mttf=mttf, mttr=mttr, load=rts79.load_curve_mw(), step=step, for_rate=np.array(q),
                                                probs=np.array(probs), lole=res.lole_hours_yr, eue=res.eue_mwh_yr)
    print("analytical", step, res.lole_hours_yr, res.eue_mwh_yr, len(probs), flush=True)
for k, v in cases.items():
    np.savez_compressed(os.path.join(out, f"ref_{k}.npz"), reference_sha256=np.array(sha), **v)
print("reference text sha256", sha)
