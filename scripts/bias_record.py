"""Bias record of the sampler paths: very long runs against the analytical COPT values.
RTS-79 sequential 1e10 years (10 seeds x 1e9), RTS-79 state sampling 1e11 samples, config 5 sequential 2e8 years.
Writes gpurun_out/bias_record.json; z = (estimate - analytical) / standard error."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from powersystemsreliabilityassessment_b200 import Engine, indices_from_raw, rts79

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
out = {}


def analytical(e, cap, mttf, mttr, load):
    lam = 1 / mttf; mu = 1 / mttr; q = lam / (lam + mu)
    p = e.copt(cap, q, 1.0)
    return e.copt_indices(p, 1.0, float(cap.sum()), load.astype(np.float64))


def seq_run(name, cap, mttf, mttr, load, seeds, years_per_seed, **kw):
    with Engine() as e:
        e.set_system(cap, mttf, mttr); e.set_load(load)
        a_lole, a_eue = analytical(e, cap, mttf, mttr, load)
        tot = None; ms = 0.0; per_seed = []
        for s in range(seeds):
            r = e.seq_mc(years_per_seed, seed=77_000 + s, **kw)
            ms += r.kernel_ms
            per_seed.append(dict(seed=77_000 + s, lole=r.lole, eens=r.eens, lolf=r.lolf))
            tot = r.raw if tot is None else {k: tot[k] + r.raw[k] for k in tot}
        idx = indices_from_raw(tot)
        out[name] = dict(years=int(tot["years"]), kernel_ms=ms, lole=idx.lole, lole_se=idx.lole_se, eens=idx.eens, eens_se=idx.eens_se,
                         lolf=idx.lolf, lold=idx.lold, analytical_lole=a_lole, analytical_eens=a_eue,
                         z_lole=(idx.lole - a_lole) / idx.lole_se, z_eens=(idx.eens - a_eue) / idx.eens_se, per_seed=per_seed, options=kw)
        print(name, json.dumps({k: v for k, v in out[name].items() if k != "per_seed"}), flush=True)


cap, mttf, mttr = rts79.units()
load = rts79.load_curve_int()
seq_run("rts79_sequential_stationary", cap, mttf, mttr, load, 10, int(1e9 * scale))
seq_run("rts79_sequential_reference_semantics_all_up_chains_of_1000_years", cap, mttf, mttr, load, 2, int(5e8 * scale), init_mode=0, years_per_chain=1000)
c5 = rts79.synthetic_system(32, 37.0)
seq_run("config5_sequential", c5[0], c5[1], c5[2], c5[3], 2, int(1e8 * scale))
with Engine() as e:
    e.set_system(cap, mttf, mttr); e.set_load(load)
    a_lole, a_eue = analytical(e, cap, mttf, mttr, load)
    n = int(1e11 * scale); tot = None; ms = 0.0
    for s in range(10):
        g = e.nonseq_mc(n // 10, seed=88_000 + s)
        ms += g["kernel_ms"]
        tot = g["raw"] if tot is None else {k: tot[k] + g["raw"][k] for k in tot}
    N = tot["samples"]
    ml = tot["sum_lol_hours"] / N; me = tot["sum_ens_fp"] / N
    sl = (max(tot["sum_lol_sq"] / N - ml * ml, 0.0) / N) ** 0.5; se = (max(tot["sum_ens_sq"] / N - me * me, 0.0) / N) ** 0.5
    out["rts79_state_sampling"] = dict(samples=int(N), kernel_ms=ms, lole=ml, lole_se=sl, eue=me, eue_se=se, analytical_lole=a_lole,
                                       analytical_eue=a_eue, z_lole=(ml - a_lole) / sl, z_eue=(me - a_eue) / se)
    print("rts79_state_sampling", json.dumps(out["rts79_state_sampling"]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bias_record.json", "w"), indent=1)
