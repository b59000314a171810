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

# ====================================================================== the "next" rows (other reference files)
from oracle import oracle as O

sha_ma = hashlib.sha256(J.load_text(ref_root, J.MULTI_AREA_REL).encode()).hexdigest()
src_ma = J.load_text(ref_root, J.MULTI_AREA_REL)
src_tail = J.load_text(ref_root, J.TAIL_RISK_REL)
src_comp = J.load_text(ref_root, J.COMPREHENSIVE_REL)
more = {}

# --- solve_curtailment_fast: random margins / topologies, both policies
rng = np.random.default_rng(4242)
topos, margins, pols, curts = [], [], [], []
for t in range(400):
    n = int(rng.integers(2, 7))
    topo = np.zeros((6, 6))
    for i in range(n):
        for j in range(i + 1, n):
            if rng.random() < 0.5:
                c = float(rng.integers(10, 300)); topo[i, j] += c; topo[j, i] += c
    m = np.zeros(6); m[:n] = rng.integers(-400, 400, n)
    pol = int(t & 1)
    c = np.zeros(6); c[:n] = J.reference_solve_curtailment(src_ma, topo[:n, :n], m[:n], pol)
    topos.append(topo); margins.append(m); pols.append(pol); curts.append(c)
more["solve_curtailment"] = dict(n_areas=np.array([int((np.abs(m) > 0).sum() and len(m)) for m in margins]), topology=np.array(topos),
                                 margins=np.array(margins), policy=np.array(pols), curtailment=np.array(curts))
more["solve_curtailment"]["n_areas"] = np.array([max(2, int(np.max(np.nonzero(np.abs(t).sum(0) + np.abs(m))[0]) + 1) if (np.abs(t).sum() + np.abs(m).sum()) > 0 else 2)
                                                 for t, m in zip(topos, margins)])

# --- run_fast_sequential_simulation, one all-up year at a time, durations = the library's sampler streams of (seed; year, unit)
def sampler_lists(seed, year, mttf, mttr, K):
    """D[u][k]: k = 0 the initial time to failure (draw 1 of the stream; draw 0 is the initial-state draw), then repair, failure, ..."""
    D = np.empty((len(mttf), K))
    for u in range(len(mttf)):
        words = []
        for b in range((K + 1 + 3) // 4):
            words.extend(int(w) for w in O.philox([year & 0xffffffff, year >> 32, u, b], [seed & 0xffffffff, seed >> 32]))
        for k in range(K):
            D[u, k] = O.duration_hours(mttf[u] if k % 2 == 0 else mttr[u], words[k + 1])
    return D


xx = np.linspace(0.0, 2.0 * np.pi, 8760)
systems = {
    "demo2": dict(unit_area=np.array([0] * 5 + [1] * 5), cap=np.array([400.0] * 5 + [200.0] * 5), mttf=np.array([1000.0] * 5 + [900.0] * 5),
                  mttr=np.array([50.0] * 5 + [60.0] * 5), loads=np.stack([np.rint(1000.0 + 500.0 * np.sin(xx)), np.rint(800.0 + 400.0 * np.sin(xx))]),
                  topology=np.array([[0.0, 200.0], [200.0, 0.0]])),
    "mesh3": dict(unit_area=np.array([0] * 4 + [1] * 3 + [2] * 4), cap=np.array([300.0, 300, 200, 100, 250, 250, 150, 200, 200, 100, 50]),
                  mttf=np.array([800.0, 900, 1100, 600, 1000, 700, 500, 950, 850, 400, 300]), mttr=np.array([60.0, 50, 40, 30, 70, 45, 25, 55, 65, 20, 15]),
                  loads=np.stack([np.rint(620.0 + 200.0 * np.sin(xx)), np.rint(450.0 + 150.0 * np.cos(xx)), np.rint(400.0 + 120.0 * np.sin(2 * xx))]),
                  topology=np.array([[0.0, 80.0, 60.0], [80.0, 0.0, 40.0], [60.0, 40.0, 0.0]])),
}
for name, sy in systems.items():
    seed, year0, ny = 2026, 3, 4
    out_l = np.zeros((2, ny, len(sy["loads"]))); out_e = np.zeros_like(out_l)
    for pol in (0, 1):
        for y in range(ny):
            D = sampler_lists(seed, year0 + y, sy["mttf"], sy["mttr"], 96)
            l, e, hit = J.reference_multi_area_year(src_ma, sy["unit_area"], sy["cap"], sy["mttf"], sy["mttr"], sy["loads"], sy["topology"], pol, D)
            assert hit == [210, 213], hit
            out_l[pol, y] = l; out_e[pol, y] = e
    more["multi_area_" + name] = dict(seed=seed, year0=year0, lole=out_l, eue=out_e, **sy)
    print("multi-area", name, out_l.sum(axis=1), flush=True)

# --- schedule_maintenance! + run_detailed_mc on the 6-unit system of tail_risk.jl, 3 years, recorded rand() / randn()
cap6 = np.array([400.0, 300.0, 300.0, 150.0, 200.0, 56.0]); q6 = np.array([0.02, 0.04, 0.04, 0.05, 0.01, 0.10])
mw6 = np.array([4, 3, 3, 2, 2, 0]); el6 = np.array([np.inf, np.inf, np.inf, np.inf, 200.0 * 50.0, np.inf])
rng = np.random.default_rng(77)
hh = np.arange(1, 8761)
base = np.maximum(0.0, 750.0 + 300.0 * np.sin((hh - 2000) / 8760 * 2 * np.pi) + 50.0 * rng.standard_normal(8760))
peaks = [base[(w - 1) * 168:min(w * 168, 8760)].max() for w in range(1, 53)]
ms6 = np.array(J.reference_schedule_maintenance(src_comp, cap6, mw6, peaks))
unif = rng.random((3, 8760, 6)); norm = rng.standard_normal((3, 8760))
yl, hf = J.reference_detailed_mc(src_tail, cap6, q6, ms6, mw6, el6, base, 5.0, 3, unif, norm)
more["detailed_mc"] = dict(cap=cap6, for_rate=q6, maint_weeks=mw6, maint_start=ms6, energy_limit=el6, base_load=base, weekly_peaks=np.array(peaks),
                           lfu_sigma_percent=5.0, unif=unif.astype(np.float32).astype(np.float64), norm=norm, yearly_lole=np.array(yl), hourly_failure_prob=np.array(hf))
# (the uniforms are stored with binary32 precision to keep the fixture small; the run above must use the stored values)
unif = more["detailed_mc"]["unif"]
yl, hf = J.reference_detailed_mc(src_tail, cap6, q6, ms6, mw6, el6, base, 5.0, 3, unif, norm)
more["detailed_mc"]["yearly_lole"] = np.array(yl); more["detailed_mc"]["hourly_failure_prob"] = np.array(hf)
print("detailed MC", yl, "maintenance starts", ms6, flush=True)
for k, v in more.items():
    np.savez_compressed(os.path.join(out, f"ref_{k}.npz"), **v)

# --- frequency & duration recursion and the stand-alone COPT demo
src_fd = J.load_text(ref_root, J.FREQUENCY_REL); src_gaa = J.load_text(ref_root, J.ASSESSMENT_REL)
rng = np.random.default_rng(5150)
fd_cases = [(np.array([16.0, 16.0]), np.array([4380.0, 4380.0]), np.array([89.39, 89.39]), 20.0)]        # the file's own demo (:199-216)
for _ in range(3):
    U = int(rng.integers(3, 7)); c = rng.integers(5, 60, U).astype(np.float64)
    fd_cases.append((c, rng.uniform(800.0, 5000.0, U), rng.uniform(20.0, 200.0, U), float(np.rint(0.7 * c.sum()))))
fd = {}
for k, (c, a, b, peak) in enumerate(fd_cases):
    lv, Pc, Fc, risk = J.reference_fd(src_fd, c, a, b, peak)
    fd.update({f"cap{k}": c, f"mtbf{k}": a, f"mttr{k}": b, f"peak{k}": peak, f"P{k}": np.array(Pc), f"F{k}": np.array(Fc), f"risk{k}": np.array(risk)})
    print("F&D", k, risk, flush=True)
np.savez_compressed(os.path.join(out, "ref_fd.npz"), n=len(fd_cases), **fd)
gaa = {}
gaa_cases = [(np.array([50.0, 50.0, 100.0, 100.0, 200.0]), np.array([0.02, 0.02, 0.03, 0.03, 0.05]), 10.0),
             (np.array([40.0, 45.0, 75.0, 120.0, 33.0, 12.0]), np.array([0.015, 0.02, 0.04, 0.06, 0.08, 0.1]), 7.0),
             (np.array([25.0, 25.0, 60.0, 110.0]), np.array([0.02, 0.02, 0.03, 0.04]), 5.0)]
for k, (c, q, step) in enumerate(gaa_cases):
    ldc = np.sort(rng.uniform(0.35, 0.95, 300) * c.sum())[::-1].copy()
    probs, idx = J.reference_gaa(src_gaa, c, q, step, ldc)
    gaa.update({f"cap{k}": c, f"q{k}": q, f"step{k}": step, f"ldc{k}": ldc, f"probs{k}": np.array(probs), f"idx{k}": np.array(idx)})
    print("GAA", k, idx, flush=True)
np.savez_compressed(os.path.join(out, "ref_gaa.npz"), n=len(gaa_cases), **gaa)

# --- Markov_process.jl (a script): the constant-hazard experiment (:46-60) and the five-generator DTMC (:153-195)
src_mk = J.load_text(ref_root, J.MARKOV_REL)
rng = np.random.default_rng(808)
u_dt = rng.random((1000, 5)).astype(np.float32).astype(np.float64)
series, mf, mr, cp = J.reference_dtmc_capacity(src_mk, u_dt)
lam_ft = 1.0 / 2500.0
u_ft = rng.random((40, 5002)).astype(np.float32).astype(np.float64)
ft, _ = J.reference_failure_times(src_mk, lam_ft, 1.0, 5000, 40, u_ft)
# the two-state chain of PART 4 (:83-110): the script's own MTTF / MTTR / dt / 200 steps, and a second parameter set
m2a, lam2, mu2, dt2, m2_hit = J.reference_markov2(src_mk)
assert m2_hit == [94, 102, 107], m2_hit
m2b, _, _, _, _ = J.reference_markov2(src_mk, 450.0, 20.0, 500)
np.savez_compressed(os.path.join(out, "ref_markov.npz"), dtmc_uniforms=u_dt, dtmc_capacity=np.array(series), mttf=np.array(mf), mttr=np.array(mr),
                    cap=np.array(cp), ft_lambda=lam_ft, ft_uniforms=u_ft, failure_times=np.array(ft),
                    markov2_script=np.array(m2a), markov2_script_params=np.array([1.0 / lam2, 1.0 / mu2, dt2]), markov2_b=np.array(m2b),
                    markov2_b_params=np.array([450.0, 20.0, 1.0]))
print("Markov: DTMC mean capacity", np.mean(series), "failure times", len(ft), "of 40", flush=True)

# --- run_detailed_analytical (tail_risk.jl:96-141) with update_elu! / calculate_expected_generation / add_unit of comprehensive.jl
d = np.load(os.path.join(out, "ref_detailed_mc.npz"))
import time as _t
_t0 = _t.time()
tot, prof, qeff, hist = J.reference_detailed_analytical(src_tail, src_comp, d["cap"], d["for_rate"], d["maint_start"], d["maint_weeks"],
                                                        d["energy_limit"], d["base_load"], 5.0)
np.savez_compressed(os.path.join(out, "ref_detailed_analytical.npz"), total=tot, profile=np.array(prof), effective_q=np.array(qeff),
                    history_q_elu=np.array(hist[4]))
print("detailed analytical: total risk", tot, "effective q", qeff, "ELU history", hist[4], f"({_t.time() - _t0:.0f} s)", flush=True)

# ====================================================================== the MATLAB functions on the path (SURVEY a-8, a-9)
# seq_mcsampling.m (next-event sampler, round / ceil discretisation, all components UP at hour 0 of every year:
# seqMain.m:91 calls it with num_years = 1) and calnlc.m, transliterated by oracle/m_transliterate.py.  The two duration draws
# (:52,59) read the library's sampler durations of the (seed; year, unit) streams; the state matrix is evaluated at HL1
# (available capacity against the load, strict compare as PSA.jl:253) and calnlc counts the curtailment events.
from oracle import m_transliterate as MT

m_sampler, _, m_hit = MT.load_seq_mcsampling(ref_root)
m_calnlc, _ = MT.load_calnlc(ref_root)
assert m_hit == [52, 59], m_hit
sha_m = hashlib.sha256((MT._load(ref_root, MT.SAMPLING_REL) + MT._load(ref_root, MT.CALNLC_REL)).encode("utf-8")).hexdigest()
rng = np.random.default_rng(9119)
m_cases = {
    "rts79": dict(cap=cap, mttf=mttf, mttr=mttr, load=rts79.load_curve_int().astype(np.float64), seed=77, year0=32, years=12),
    # frequent events on a ragged horizon: times to failure that round to 0 hours, repairs that end beyond the year
    "small": dict(cap=rng.integers(5, 60, 5).astype(np.float64), mttf=rng.uniform(6.0, 40.0, 5), mttr=rng.uniform(0.6, 9.0, 5),
                  load=rng.integers(60, 150, 301).astype(np.float64), seed=5, year0=0, years=40),
}
mat = {}
for name, c in m_cases.items():
    U, H = len(c["cap"]), len(c["load"])
    mf32 = c["mttf"].astype(np.float32).astype(np.float64); mr32 = c["mttr"].astype(np.float32).astype(np.float64)   # the sampler's binary32 means
    lol = np.zeros(c["years"]); ens = np.zeros(c["years"]); nlc = np.zeros(c["years"]); down = np.zeros((c["years"], U))
    for y in range(c["years"]):
        D = sampler_lists(c["seed"], c["year0"] + y, mf32, mr32, 400 if name == "small" else 96)
        states, used = m_sampler(np.stack([mf32, mr32], axis=1), U, 0, 1, H, D)
        st = np.array(states)
        cap_avail = ((1.0 - st) * c["cap"][:, None]).sum(axis=0)        # whole-MW capacities: the sum is exact in any order
        flag = cap_avail < c["load"]
        lol[y] = flag.sum(); ens[y] = (c["load"] - cap_avail)[flag].sum(); nlc[y] = m_calnlc(flag.astype(np.float64)); down[y] = st.sum(axis=1)
    mat.update({f"{name}_{k}": v for k, v in c.items()})
    mat.update({f"{name}_lol": lol, f"{name}_ens": ens, f"{name}_nlc": nlc, f"{name}_down_hours": down})
    print("MATLAB sampler", name, "DLC", lol.sum(), "ENS", ens.sum(), "NLC", nlc.sum(), "down hours", down.sum(), flush=True)
np.savez_compressed(os.path.join(out, "ref_matlab.npz"), reference_sha256=np.array(sha_m), **mat)
