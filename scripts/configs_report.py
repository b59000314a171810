"""Measured results of the five BASELINE.json configs on one B200 -> gpurun_out/configs.json (copied to profiles/)."""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import powersystemsreliabilityassessment_b200 as P
from powersystemsreliabilityassessment_b200 import rts79

out = {}
cap, mttf, mttr = rts79.units()
load = rts79.load_curve_int()
lam = 1 / mttf; mu = 1 / mttr; q = lam / (lam + mu)
with P.Engine() as e:
    # C1: non-sequential, 1e5 samples, hourly curve and annual peak
    e.set_system(cap, mttf, mttr); e.set_load(load)
    e.nonseq_mc(1000, seed=1)
    g = e.nonseq_mc(100_000, seed=42)
    e.set_load(np.array([2850], dtype=np.int32))
    p = e.nonseq_mc(100_000, seed=42)
    big = None
    e.set_load(load)
    big = e.nonseq_mc(100_000_000, seed=7)
    out["C1_nonseq_1e5"] = dict(lole=g["lole"], lole_se=g["lole_se"], eue=g["eue"], eue_se=g["eue_se"], kernel_ms=g["kernel_ms"],
                                peak_plc=p["lole"], peak_plc_se=p["lole_se"], peak_edns_mw=p["eue"], peak_lole_h=p["lole"] * 8760,
                                analytical_lole=9.3677375218, analytical_plc=0.084578060826,
                                throughput_samples_per_s_at_1e8=1e8 / big["kernel_ms"] * 1e3)
    # C2: sequential, 1e4 years
    e.seq_mc(1000, seed=1)
    r = e.seq_mc(10_000, seed=42, group=10)
    out["C2_seq_1e4"] = dict(lole=r.lole, lole_se=r.lole_se, eens=r.eens, eens_se=r.eens_se, lolf=r.lolf, lold=r.lold,
                             p_loss_year=r.p_loss_year, kernel_ms=r.kernel_ms, years_per_s=1e4 / r.kernel_ms * 1e3)
    # C3: analytical COPT / F&D / Markov
    t0 = time.perf_counter(); pr = e.copt(cap, q, 1.0); l1, e1 = e.copt_indices(pr, 1.0, 3405.0, rts79.load_curve_mw()); t1 = time.perf_counter()
    pr10 = e.copt(cap, q, 10.0); l10, e10 = e.copt_indices(pr10, 10.0, 3405.0, rts79.load_curve_mw())
    Pc, Fc = e.fd_recursion(cap, mttf + mttr, mttr)
    lole_fd, lolf_fd, lold_fd = P.evaluate_risk(Pc, Fc, 2850.0, 3405.0)
    m = e.markov2(1000.0, 50.0, 1.0, 200)
    out["C3_analytical"] = dict(copt_step1=dict(states=len(pr), lole=l1, eue=e1, wall_ms=(t1 - t0) * 1e3),
                                copt_step10=dict(states=len(pr10), lole=l10, eue=e10),
                                fd_at_peak=dict(lole_h=lole_fd, lolf=lolf_fd, lold=lold_fd),
                                markov_pdown_200=float(m[-1]))
    # C4: tail risk over 1e6 years
    r = e.seq_mc(1_000_000, seed=42, keep_on_device=True)
    t0 = time.perf_counter(); tail, hist = e.tail(None, alphas=(0.95, 0.99), n_bins=50, bin_width=1000); t1 = time.perf_counter()
    out["C4_tail_1e6"] = dict(seq_kernel_ms=r.kernel_ms, tail_wall_ms=(t1 - t0) * 1e3, var95=tail[0]["var"], cvar95=tail[0]["cvar"],
                              var99=tail[1]["var"], cvar99=tail[1]["cvar"], lole=r.lole, eens=r.eens, hist_first_bins=[int(x) for x in hist[:6]])
    # weak-point detection (seqMain.m:225-231 at HL1) over 1e6 years: generic kernel + replay of the loss segments
    imp, cnt, ri = e.seq_unit_importance(1_000_000, seed=42)
    out["unit_importance_1e6"] = dict(kernel_ms=ri.kernel_ms, years_per_s=1e6 / ri.kernel_ms * 1e3, lole=ri.lole,
                                      top5=[(int(u), float(imp[u])) for u in np.argsort(-imp)[:5]])
    # detailed MC (tail_risk.jl engine), 2000 years and 1e5 years
    gens = [P.DetailedGenerator("Nuclear", 400.0, 0.02, 4), P.DetailedGenerator("Coal_A", 300.0, 0.04, 3),
            P.DetailedGenerator("Coal_B", 300.0, 0.04, 3), P.DetailedGenerator("Gas", 150.0, 0.05, 2),
            P.DetailedGenerator("Hydro_ELU", 200.0, 0.01, 2, 200.0 * 50.0), P.DetailedGenerator("Old_56", 56.0, 0.10, 0)]
    rng = np.random.default_rng(7); h = np.arange(1, 8761)
    base = np.maximum(0.0, 750.0 + 300.0 * np.sin((h - 2000) / 8760 * 2 * math.pi) + 50.0 * rng.standard_normal(8760))
    P.schedule_maintenance(gens, [base[(w - 1) * 168:min(w * 168, 8760)].max() for w in range(1, 53)])
    yl, hf, ms = e.detailed_mc(gens, base, base.max() * 0.05, 100_000, seed=1)
    out["detailed_mc_1e5"] = dict(kernel_ms=ms, years_per_s=1e5 / ms * 1e3, mean_lole=float(yl.mean()), p95=float(np.quantile(yl, 0.95)))
    yl, hf, ms = e.detailed_mc(gens, base, base.max() * 0.05, 2_000_000, seed=1)          # thread = year: the GPU fills up at ~3e5 years
    out["detailed_mc_2e6"] = dict(kernel_ms=ms, years_per_s=2e6 / ms * 1e3, hour_steps_per_s=2e6 * 8760 / ms * 1e3, mean_lole=float(yl.mean()))
    # multi-area (AdequacyAssessmentII.jl demo: 2 areas x 5 units, one 200 MW tie), both policies
    ua = np.array([0] * 5 + [1] * 5); acap = np.array([400.0] * 5 + [200.0] * 5)
    amttf = np.array([1000.0] * 5 + [900.0] * 5); amttr = np.array([50.0] * 5 + [60.0] * 5)
    xx = np.linspace(0.0, 2.0 * np.pi, 8760)
    loads = np.stack([np.rint(1000.0 + 500.0 * np.sin(xx)), np.rint(800.0 + 400.0 * np.sin(xx))])
    topo = np.array([[0.0, 200.0], [200.0, 0.0]])
    for pol, name in ((0, "isolated"), (1, "interconnected")):
        e.multi_area_mc(ua, acap, amttf, amttr, loads, topo, pol, 2000, seed=6)
        m = e.multi_area_mc(ua, acap, amttf, amttr, loads, topo, pol, 1_000_000, seed=6)
        out[f"multi_area_demo_{name}_1e6"] = dict(kernel_ms=m["kernel_ms"], years_per_s=1e6 / m["kernel_ms"] * 1e3,
                                                  lole=[float(v) for v in np.atleast_1d(m["lole"])], eue=[float(v) for v in np.atleast_1d(m["eue"])])
    # C5: 1024 units
    c5 = rts79.synthetic_system(32, 37.0)
    e.set_system(c5[0], c5[1], c5[2]); e.set_load(c5[3])
    e.seq_mc(2000, seed=1)
    r = e.seq_mc(1_000_000, seed=42)
    lam5 = 1 / c5[1]; mu5 = 1 / c5[2]; q5 = lam5 / (lam5 + mu5)
    p5 = e.copt(c5[0], q5, 1.0); l5, e5 = e.copt_indices(p5, 1.0, float(c5[0].sum()), c5[3].astype(float))
    out["C5_1024_units"] = dict(years=1_000_000, kernel_ms=r.kernel_ms, years_per_s=1e6 / r.kernel_ms * 1e3, lole=r.lole, lole_se=r.lole_se,
                                eens=r.eens, eens_se=r.eens_se, lolf=r.lolf, analytical_lole=l5, analytical_eue=e5, copt_states=len(p5))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)
print(json.dumps(out, indent=1))
