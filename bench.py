#!/usr/bin/env python
"""bench.py -- simulated system-years/s of the RTS-79 sequential HL1 Monte Carlo on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, libpsra_b200.so)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU restatement arm

A "step" = one pass of the hot path (psra_seq_mc) over --years-per-step RTS-79 system-years on
every GPU (weak scaling: ranks own disjoint contiguous year ranges of one experiment, the integer
accumulators are summed with one NCCL all-reduce per step).  `value` = years of all ranks / time with
the system data already resident in HBM; `e2e` = the same through the reference-facing API
run_sequential_mc (host system data uploaded, result + convergence history read back, every step) --
at N > 1 that is ONE call from rank 0 on an engine with ngpus = N: the library shards the years over the
N devices and all-reduces the integers with NCCL itself (csrc/multi.cu); the other ranks wait on the host.
`configs` carries the other BASELINE configurations measured in the same run: config 5 (1024 units) and
RTS-79 as STRONG scaling (a fixed 10^7-year experiment split over the N ranks), and on rank 0 config 1
(state sampling), config 3 (COPT) and config 4 (10^6 years + VaR / CVaR from the in-kernel histogram).
Contract details: module docstring of the task / DESIGN.md section 6.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "simulated system-years/sec (RTS-79 sequential HL1)"
UNIT = "system-years/s"
WORKLOAD = ("IEEE RTS-79 HL1 sequential chronological MCS, 32 units, 8736-h integer-MW load curve, "
            "exponential TTF/TTR, Philox4x32-10 keyed (seed; year, unit), STATIONARY start, "
            "LOLE/EENS/LOLF/duration accumulators")


DTYPE = "int32 MW / int64 ticks of 2^-24 h + binary32 logarithm (no FP64 on the device path)"


def w_alg_thread_instr(mttf, mttr, H):
    """SURVEY.md 8d algorithmic work per system-year: 8 per hour slot + 30 per RNG event."""
    events = len(mttf) + float((2.0 * H / (mttf + mttr)).sum())
    return 8.0 * H + 30.0 * events, events


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(budget_s: float, threads: int):
    """The oracle's literal PSA.jl:214-269 hour/unit loop (sampler-driven, gcc -O3), trial-parallel over host
    threads (ctypes releases the GIL).  Returns (years/s, years simulated, seconds)."""
    from oracle import oracle as O
    from powersystemsreliabilityassessment_b200 import rts79
    cap, mttf, mttr = rts79.units()
    load = rts79.load_curve_int().astype(np.float64)
    O.lib()
    t0 = time.perf_counter()
    O.seq_philox(cap, mttf, mttr, load, 1, 0, 100, 1, 1)
    per_year = (time.perf_counter() - t0) / 100
    n_each = max(50, int(budget_s / per_year))
    done = [0] * threads

    def work(i):
        O.seq_philox(cap, mttf, mttr, load, 42, i * n_each, n_each, 1, 1)
        done[i] = n_each

    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return sum(done) / dt, sum(done), dt


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; Julia is not installed on the box)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from oracle import oracle as O
    from powersystemsreliabilityassessment_b200 import rts79
    cap, mttf, mttr = rts79.units()
    load = rts79.load_curve_int().astype(np.float64)
    O.lib()
    t0 = time.perf_counter()
    O.seq_philox(cap, mttf, mttr, load, 1, 0, 64, 1, 1)
    per_year = (time.perf_counter() - t0) / 64
    total_budget = args.ref_budget                           # whole run bounded (default ~1.5 min of wall time)
    n_each = max(16, int(total_budget / (args.steps + args.warmup) / per_year))

    def step(s):
        ths = [threading.Thread(target=O.seq_philox, args=(cap, mttf, mttr, load, 42, (s * threads + i) * n_each,
                                                           n_each, 1, 1)) for i in range(threads)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    for s in range(args.warmup):
        step(s)
    t0 = time.perf_counter()
    for s in range(args.steps):
        step(args.warmup + s)
    dt = time.perf_counter() - t0
    years = n_each * threads * args.steps
    val = years / dt
    sample = f"{n_each * threads} system-years per step ({n_each} per thread x {threads} threads), literal hour/unit loop"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "years_per_step": n_each * threads, "host_threads": threads,
                   "note": "the metric is a rate: the CPU arm simulates a bounded sample of the same workload per step "
                           "(the GPU arm's 10^7 years per step would take ~5 min per step here)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement of PowerSystemAdequacy.jl:214-269 (oracle/psra_oracle.c, gcc -O3); Julia/MATLAB are not "
                "installed, so the reference scripts themselves cannot be timed",
    }))


def bench_configs(P, rts79, sharding, eng, dev, rank, world, seed, sm_count, sm_max, hostbar, torch, dist):
    """The BASELINE configurations beside the headline, in the same run.  c5 / rts79_strong: a FIXED 10^7-year
    experiment split over the ranks (strong scaling; device time = max over ranks, CUDA events on the library's
    stream); c1 / c3 / c4 on rank 0."""
    from powersystemsreliabilityassessment_b200 import sharding as S
    ext = torch.cuda.ExternalStream(eng.stream, device=dev)
    peak = sm_count * 4 * sm_max * 1e6
    out = {}

    def strong(total_years, label):
        y0, y1 = S.shard_range(total_years, rank, world)
        eng.seq_mc(max(1, (y1 - y0) // 50), seed=seed + 1, year0=y0)          # warm-up (kernel attributes, clocks)
        torch.cuda.synchronize(dev)
        hostbar()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(ext)
        r = eng.seq_mc(y1 - y0, seed=seed, year0=y0)
        red = S.allreduce_raw(r.raw, device=dev)
        e1.record(ext)
        torch.cuda.synchronize(dev)
        hostbar()
        wall = time.perf_counter() - t0
        t = torch.tensor([e0.elapsed_time(e1), r.kernel_ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, k_ms, wall_ms = t.tolist()
        idx = P.indices_from_raw(red)
        return dict(scaling="strong", years_total=total_years, n_gpus=world, device_ms=dev_ms, kernel_ms_max=k_ms, wall_ms=wall_ms,
                    years_per_s=total_years / (dev_ms * 1e-3), lole=idx.lole, lole_se=idx.lole_se, eens=idx.eens,
                    eens_se=idx.eens_se, lolf=idx.lolf, redone=int(r.redone)), k_ms

    # ---- config 5: 1024 units = RTS-79 x 32, load x 37 (analytical LOLE 8.0331 h/yr), 10^7 years in total
    cap5, mttf5, mttr5, load5 = rts79.synthetic_system(32, 37.0)
    eng.set_system(cap5, mttf5, mttr5); eng.set_load(load5)
    c5, k_ms = strong(10_000_000, "c5")
    w5, ev5 = w_alg_thread_instr(mttf5, mttr5, len(load5))
    y_local = S.shard_range(10_000_000, 0, world)
    y_local = y_local[1] - y_local[0]
    c5.update(units=len(cap5), hours=len(load5), analytical_lole=8.0331, alg_warp_inst_per_year=w5 / 32.0, events_per_year=ev5,
              kernel="seq_wide_kernel (csrc/seq_wide.cu)",
              roofline_frac=(w5 / 32.0) * y_local / (k_ms * 1e-3) / peak,
              z_lole=(c5["lole"] - 8.0331) / c5["lole_se"] if c5["lole_se"] > 0 else None)
    out["c5"] = c5

    # ---- RTS-79, 10^7 years in total: time to solution incl. launch, all-reduce and host latency
    cap, mttf, mttr = rts79.units()
    load = rts79.load_curve_int()
    eng.set_system(cap, mttf, mttr); eng.set_load(load)
    rs, _ = strong(10_000_000, "rts79")
    rs.update(analytical_lole=9.3677375218)
    out["rts79_strong"] = rs

    hostbar()
    if rank != 0:
        return out

    # ---- config 1: non-sequential state sampling, 1e5 samples (the BASELINE size) and 1e8 (throughput)
    torch.cuda.synchronize(dev)
    eng.nonseq_mc(100_000, seed=seed)
    r5 = eng.nonseq_mc(100_000, seed=seed)
    r8 = eng.nonseq_mc(100_000_000, seed=seed)
    w1 = (20.0 * len(cap) + 30.0) / 32.0
    out["c1"] = dict(samples=100_000, kernel_ms=r5["kernel_ms"], lole=r5["lole"], lole_se=r5["lole_se"], eue=r5["eue"],
                     samples_1e8_kernel_ms=r8["kernel_ms"], samples_per_s=1e8 / (r8["kernel_ms"] * 1e-3),
                     alg_warp_inst_per_sample=w1, roofline_frac=w1 * 1e8 / (r8["kernel_ms"] * 1e-3) / peak,
                     lole_1e8=r8["lole"], lole_1e8_se=r8["lole_se"], analytical_lole=9.3677375218,
                     kernel="nonseq_fast_kernel (csrc/nonseq_mc.cu)")
    eng.set_load(np.array([2850]))
    pk = eng.nonseq_mc(100_000, seed=seed)
    out["c1"]["peak_load_plc"] = pk["p_loss"]; out["c1"]["peak_load_lole_8760"] = pk["p_loss"] * 8760.0
    eng.set_load(load)

    # ---- config 3: COPT (1 MW and 10 MW grids) + indices, wall time of the API calls
    lam = 1.0 / mttf; mu = 1.0 / mttr; q = lam / (lam + mu)
    gens = [P.Generator(i + 1, float(c), float(a), float(b)) for i, (c, a, b) in enumerate(zip(cap, mttf, mttr))]
    lm = P.LoadModel(rts79.load_curve_mw())
    c3 = {}
    for step in (1.0, 10.0):
        P.run_analytical(gens, lm, step_size=step, engine=eng)
        t0 = time.perf_counter()
        ra = P.run_analytical(gens, lm, step_size=step, engine=eng)
        c3[f"step{int(step)}"] = dict(wall_ms=(time.perf_counter() - t0) * 1e3, lole=ra.lole_hours_yr, eue=ra.eue_mwh_yr,
                                      states=len(eng.copt(cap, q, step)))
    c3["known_answers"] = {"step1": [9.3941103566, 1176.291677], "step10": [9.4204746080, 1177.243237]}
    out["c3"] = c3

    # ---- config 4: 10^6 years, VaR / CVaR at 95 / 99 % from the in-kernel ENS histogram (no per-year vector)
    eng.seq_mc(100_000, seed=seed, tail_hist=True); eng.tail(None)
    t0 = time.perf_counter()
    r4 = eng.seq_mc(1_000_000, seed=seed, tail_hist=True)
    t1 = time.perf_counter()
    tail = eng.tail(None, alphas=(0.95, 0.99))
    t2 = time.perf_counter()
    r4p = eng.seq_mc(1_000_000, seed=seed)
    out["c4"] = dict(years=1_000_000, seq_kernel_ms=r4.kernel_ms, seq_kernel_ms_without_histogram=r4p.kernel_ms,
                     seq_wall_ms=(t1 - t0) * 1e3, tail_wall_ms=(t2 - t1) * 1e3,
                     var95_mwh=tail[0]["var"], cvar95_mwh=tail[0]["cvar"], var99_mwh=tail[1]["var"], cvar99_mwh=tail[1]["cvar"],
                     per_year_vectors_in_hbm=0, method="1 MWh bins counted in the MC kernel; one launch over the bins (csrc/tail.cu)")

    # ---- the rows SURVEY.md section 8 marks "next": hourly-resampled detailed MC (tail_risk.jl:12-91, its 6-unit system) and
    #      the multi-area simulation (AdequacyAssessmentII.jl demo: 2 areas x 5 units, one 200 MW tie), kernel times
    import math
    dg = [P.DetailedGenerator("Nuclear", 400.0, 0.02, 4), P.DetailedGenerator("Coal_A", 300.0, 0.04, 3),
          P.DetailedGenerator("Coal_B", 300.0, 0.04, 3), P.DetailedGenerator("Gas", 150.0, 0.05, 2),
          P.DetailedGenerator("Hydro_ELU", 200.0, 0.01, 2, 200.0 * 50.0), P.DetailedGenerator("Old_56", 56.0, 0.10, 0)]
    rng = np.random.default_rng(7); hh = np.arange(1, 8761)
    base = np.maximum(0.0, 750.0 + 300.0 * np.sin((hh - 2000) / 8760 * 2 * math.pi) + 50.0 * rng.standard_normal(8760))
    P.schedule_maintenance(dg, [base[(w - 1) * 168:min(w * 168, 8760)].max() for w in range(1, 53)])
    eng.detailed_mc(dg, base, base.max() * 0.05, 20_000, seed=seed)
    yl, _, ms = eng.detailed_mc(dg, base, base.max() * 0.05, 2_000_000, seed=seed)
    out["f1_detailed_mc"] = dict(years=2_000_000, kernel_ms=ms, years_per_s=2e6 / (ms * 1e-3), hour_steps_per_s=2e6 * 8760 / (ms * 1e-3),
                                 mean_lole=float(yl.mean()), kernel="detailed_mc_kernel (csrc/detailed_mc.cu)")
    ua = np.array([0] * 5 + [1] * 5); acap = np.array([400.0] * 5 + [200.0] * 5)
    amttf = np.array([1000.0] * 5 + [900.0] * 5); amttr = np.array([50.0] * 5 + [60.0] * 5)
    xx = np.linspace(0.0, 2.0 * np.pi, 8760)
    loads = np.stack([np.rint(1000.0 + 500.0 * np.sin(xx)), np.rint(800.0 + 400.0 * np.sin(xx))])
    topo = np.array([[0.0, 200.0], [200.0, 0.0]])
    out["f3_multi_area"] = {}
    for pol, name in ((0, "isolated"), (1, "interconnected")):
        eng.multi_area_mc(ua, acap, amttf, amttr, loads, topo, pol, 2000, seed=seed)
        m = eng.multi_area_mc(ua, acap, amttf, amttr, loads, topo, pol, 1_000_000, seed=seed)
        out["f3_multi_area"][name] = dict(years=1_000_000, kernel_ms=m["kernel_ms"], years_per_s=1e6 / (m["kernel_ms"] * 1e-3),
                                          lole=[float(v) for v in np.atleast_1d(m["lole"])], eue=[float(v) for v in np.atleast_1d(m["eue"])])
    eng.set_system(cap, mttf, mttr); eng.set_load(load)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--years-per-step", type=float, default=1e7, help="system-years per GPU per step")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (config 5 / strong scaling / c1 / c3 / c4)")
    ap.add_argument("--ref-budget", type=float, default=90.0, help="seconds of wall time for --impl reference")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import powersystemsreliabilityassessment_b200 as P
    from powersystemsreliabilityassessment_b200 import rts79, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    saved_stdout = None
    hostpg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version banner to the process's stdout at the first collective: park fd 1 on stderr until
        # the JSON line is printed, so that stdout carries that one line only
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
        hostpg = dist.new_group(backend="gloo")      # host-side waits that put no spinning kernel on the GPUs

    def hostbar():
        if world > 1:
            dist.barrier(group=hostpg)

    Y = int(args.years_per_step)
    cap, mttf, mttr = rts79.units()
    load_mw = rts79.load_curve_int().astype(np.float64)          # integer-MW curve (BASELINE config 2)
    H = len(load_mw)
    gens = [P.Generator(i + 1, float(c), float(a), float(b)) for i, (c, a, b) in enumerate(zip(cap, mttf, mttr))]
    lm = P.LoadModel(load_mw)
    eng = P.Engine(device=local_rank)
    eng.set_generators(gens, lm)                                  # resident in HBM for the `value` leg
    sm_count, _ = eng.device_info()
    ext = torch.cuda.ExternalStream(eng.stream, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def step_resident(s):
        """one step, inputs resident: kernel + 256 B accumulator read-back + all-reduce of the accumulators"""
        flush.zero_()
        torch.cuda.synchronize(dev)
        y0 = (s * world + rank) * Y
        r = eng.seq_mc(Y, seed=args.seed, year0=y0)
        red = sharding.allreduce_raw(r.raw, device=dev)
        return r, red

    for s in range(args.warmup):
        step_resident(s)

    # ---------------- timed: resident-input throughput ----------------
    clocks = ClockSampler(local_rank)
    sync_all()
    clocks.start()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record(ext)
    kernel_ms = []
    acc = None
    for s in range(args.steps):
        r, red = step_resident(args.warmup + s)
        kernel_ms.append(r.kernel_ms)
        acc = red if acc is None else {k: acc[k] + red[k] for k in acc}
    ev1.record(ext)
    sync_all()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    t = torch.tensor([dev_ms, wall * 1e3, sum(kernel_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, ksum_ms = t.tolist()
    live = eng.last_counters()            # counters the kernel itself kept during the last timed launch (this run, not a profile)
    ms_per_step = dev_ms / args.steps
    value = world * Y * args.steps / (dev_ms * 1e-3)

    # ---------------- timed: end to end through run_sequential_mc, ONE call per step ----------------
    # N = 1: the rank's own engine.  N > 1: rank 0 alone, engine with ngpus = N (the library shards the years over the
    # N devices and all-reduces with NCCL); the other ranks wait on the host (gloo), their GPUs idle.
    torch.cuda.synchronize(dev)
    hostbar()
    e2e_wall = 0.0
    e2e_info = {}
    if rank == 0:
        e2e_eng = eng if world == 1 else P.Engine(device=0, ngpus=world)
        flushes = [flush] if world == 1 else []
        if world > 1:
            for g in range(world):
                flushes.append(flush if g == local_rank else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=torch.device("cuda", g)))

        flush_s = [0.0]

        def step_e2e(s):
            tf = time.perf_counter()
            for g, f in enumerate(flushes):
                f.zero_()
            for g in range(len(flushes)):
                torch.cuda.synchronize(torch.device("cuda", g if world > 1 else local_rank))
            flush_s[0] += time.perf_counter() - tf
            return P.run_sequential_mc(gens, lm, world * Y, seed=args.seed, year0=s * world * Y, engine=e2e_eng, details=True)

        step_e2e(1000)
        flush_s[0] = 0.0
        t0 = time.perf_counter()
        for s in range(args.steps):
            res, r2 = step_e2e(2000 + s)
        e2e_wall = time.perf_counter() - t0
        e2e_info = {"lole_last_step": res.lole_hours_yr, "history_len": int(len(res.convergence_history)),
                    "years_last_step": int(r2.years), "kernel_ms_last_step": r2.kernel_ms,
                    "l2_flush_ms_per_step": 1e3 * flush_s[0] / args.steps}
        if e2e_eng is not eng:
            e2e_eng.close()
        del flushes
    hostbar()
    e2e_value = world * Y * args.steps / e2e_wall if e2e_wall > 0 else None
    U = len(cap); Wd = (H + 31) // 32
    h2d = world * (U * (4 + 4 + 4 + 4 + 8 + 4 + 16) + Wd * 32 * 4 + Wd * 4)          # psra_set_system + psra_set_load uploads, per device
    d2h = 32 * 8 + 8 * ((world * Y) // 10)                                    # accumulators + LOLE history

    sm_max_peaks = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        sm_max_peaks = peaks.get("sm_max_mhz")
    except Exception:
        pass
    sm_max = float(sm_max_peaks or clk.get("sm_max_mhz") or 1965.0)

    configs = None
    if not args.no_configs:
        configs = bench_configs(P, rts79, sharding, eng, dev, rank, world, args.seed, sm_count, sm_max, hostbar, torch, dist)
        eng.set_generators(gens, lm)

    if rank == 0:
        idx = P.indices_from_raw(acc)
        w_thread, events = w_alg_thread_instr(mttf, mttr, H)
        w_warp = w_thread / 32.0
        peak = sm_count * 4 * sm_max * 1e6 / 1e9                  # Gwarp-inst/s (4 schedulers per SM)
        k_s = (ksum_ms / args.steps) * 1e-3
        achieved = w_warp * Y / k_s / 1e9
        prof = {}
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "roofline_latest.json")))
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": {"workload": WORKLOAD, "years_per_step_per_gpu": Y, "hours": H, "units": U,
                       "semantics": "independent years from the stationary law (init_mode STATIONARY, years_per_chain 1); the "
                                    "reference is one all-up chain carried across years (PSA.jl:223-224) -- same expectation, "
                                    "see INTEGRATION.md section 4",
                       "parallelism": f"years sharded over {world} GPU(s), 1 all-reduce of 11 int64 per step",
                       "l2": "256 MiB buffer written between timed steps (inside the bracket); inputs are 36 KB"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "run_sequential_mc(gens, load, years) -> ReliabilityResult incl. convergence history; ONE call per "
                           "step" + ("" if world == 1 else f" on an engine with ngpus={world}: the library shards the years over the "
                                     "devices and combines the integers with ncclAllReduce (csrc/multi.cu)"),
                    **e2e_info},
            "gpu_launches": args.steps,
            "clocks": {"sm_mhz": clk.get("sm_mhz"), "sm_max_mhz": clk.get("sm_max_mhz"), "reasons": clk.get("reasons"),
                       "samples": clk.get("samples")},
            "roofline": {"bound": "sm_issue", "achieved": achieved, "peak": peak, "unit": "Gwarp-inst/s",
                         "frac": achieved / peak, "traffic": prof.get("dram_bytes_per_launch"),
                         "kernel": "seq_fast_kernel<disc=0, ring=0, packed=1> (csrc/seq_fast.cu)", "kernel_ms_per_launch": ksum_ms / args.steps,
                         "alg_warp_inst_per_year": w_warp, "events_per_year": events,
                         "peak_source": f"{sm_count} SMs x 4 issue/clk x {sm_max:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json)",
                         "ncu_issue_active_pct": prof.get("issue_active_pct"),
                         "ncu_warp_inst_per_year": prof.get("warp_inst_per_year"),
                         "live_counters_per_year": {"philox_blocks": live["jobs"] / Y, "generation_waves": live["waves"] / Y,
                                                    "state_transitions": live["events"] / Y, "words_resolved_hour_by_hour": live["resolved_runs"] / Y,
                                                    "note": "kept by the kernel in this run; the ncu_* fields are constants of the "
                                                            "committed profile named in traffic_source"},
                         "traffic_source": prof.get("source"),
                         "hbm": {"achieved": (prof.get("dram_bytes_per_launch") or 0.0) / k_s / 1e9, "peak": peaks.get("hbm_gbs"),
                                 "unit": "GB/s", "frac": (prof.get("dram_bytes_per_launch") or 0.0) / k_s / 1e9 / float(peaks.get("hbm_gbs") or 6547.5)},
                         "hbm_note": "HBM traffic is per-launch accumulators only; not the bound (SURVEY 8d)"},
            "results": {"years": idx.years, "lole_h_per_yr": idx.lole, "lole_se": idx.lole_se,
                        "eens_mwh_per_yr": idx.eens, "eens_se": idx.eens_se, "lolf_occ_per_yr": idx.lolf,
                        "lold_h": idx.lold, "analytical_lole": 9.3677375218, "analytical_eens": 1176.181257},
            "wall_ms_per_step": wall_ms / args.steps,
        }
        if configs is not None:
            line["configs"] = configs
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rate, yrs, secs = cpu_reference_rate(10.0, threads)
            rate1, yrs1, secs1 = cpu_reference_rate(5.0, 1)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{yrs} RTS-79 system-years ({yrs // threads} per thread), literal "
                                              f"hour/unit loop of PSA.jl:214-269 (gcc -O3), {secs:.1f} s",
                                    "one_thread": {"value": rate1, "cores": 1, "sample": f"{yrs1} system-years, {secs1:.1f} s"}}
        sys.stdout.flush()
        if saved_stdout is not None:
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
