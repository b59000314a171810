#!/usr/bin/env python
"""bench.py -- simulated system-years/s of the RTS-79 sequential HL1 Monte Carlo on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, libpsra_b200.so)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU restatement arm

A "step" = one pass of the hot path (psra_seq_mc) over --years-per-step RTS-79 system-years on
every GPU (weak scaling: ranks own disjoint contiguous year ranges of one experiment, the integer
accumulators are summed with one NCCL all-reduce per step).  `value` = years of all ranks / time with
the system data already resident in HBM; `e2e` = the same through the reference-facing API
run_sequential_mc (host system data uploaded, result + convergence history read back, every step).
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
    """The oracle's literal PSA.jl:214-269 hour/unit loop (sampler-driven), trial-parallel over host
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
        "config": {"workload": WORKLOAD, "years_per_step": n_each * threads, "host_threads": threads},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement of PowerSystemAdequacy.jl:214-269 (oracle/psra_oracle.c); Julia/MATLAB are not "
                "installed, so the reference scripts themselves cannot be timed",
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--years-per-step", type=float, default=1e7, help="system-years per GPU per step")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version banner to the process's stdout at the first collective: park fd 1 on stderr until
        # the JSON line is printed, so that stdout carries that one line only
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

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

    total_raw = {}

    def step_resident(s):
        """one step, inputs resident: kernel + 256 B accumulator read-back + all-reduce of the accumulators"""
        flush.zero_()
        torch.cuda.synchronize(dev)
        y0 = (s * world + rank) * Y
        r = eng.seq_mc(Y, seed=args.seed, year0=y0)
        red = sharding.allreduce_raw(r.raw, device=dev)
        return r, red

    def step_e2e(s):
        """one step through the reference-facing API with host buffers (upload + history read-back)"""
        flush.zero_()
        torch.cuda.synchronize(dev)
        y0 = (s * world + rank) * Y
        res, r = P.run_sequential_mc(gens, lm, Y, seed=args.seed, year0=y0, engine=eng, details=True)
        red = sharding.allreduce_raw(r.raw, device=dev)
        return res, r, red

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
    ms_per_step = dev_ms / args.steps
    value = world * Y * args.steps / (dev_ms * 1e-3)

    # ---------------- timed: end to end through run_sequential_mc ----------------
    step_e2e(1000)
    sync_all()
    t0 = time.perf_counter()
    for s in range(args.steps):
        res, r2, red2 = step_e2e(2000 + s)
    sync_all()
    e2e_wall = time.perf_counter() - t0
    t = torch.tensor([e2e_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_wall = t.item()
    e2e_value = world * Y * args.steps / e2e_wall
    U = len(cap); Wd = (H + 31) // 32
    h2d = U * (4 + 4 + 4 + 4 + 8) + Wd * 32 * 4 + Wd * 4          # psra_set_system + psra_set_load uploads
    d2h = 32 * 8 + 8 * ((Y + 9) // 10)                            # accumulators + LOLE history groups

    if rank == 0:
        idx = P.indices_from_raw(acc)
        w_thread, events = w_alg_thread_instr(mttf, mttr, H)
        w_warp = w_thread / 32.0
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_max = float(peaks.get("sm_max_mhz") or clk.get("sm_max_mhz") or 1965.0)
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
            "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "years_per_step_per_gpu": Y, "hours": H, "units": U,
                       "parallelism": f"years sharded over {world} GPU(s), 1 all-reduce of 11 int64 per step",
                       "l2": "256 MiB buffer written between timed steps (inside the bracket); inputs are 36 KB"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "run_sequential_mc(gens, load, years) -> ReliabilityResult incl. convergence history"},
            "gpu_launches": args.steps,
            "clocks": {"sm_mhz": clk.get("sm_mhz"), "sm_max_mhz": clk.get("sm_max_mhz"), "reasons": clk.get("reasons"),
                       "samples": clk.get("samples")},
            "roofline": {"bound": "sm_issue", "achieved": achieved, "peak": peak, "unit": "Gwarp-inst/s",
                         "frac": achieved / peak, "traffic": prof.get("dram_bytes_per_launch"),
                         "kernel": "seq_fast_kernel<disc=0, ring=0, packed=1> (csrc/seq_fast.cu)", "kernel_ms_per_launch": ksum_ms / args.steps,
                         "alg_warp_inst_per_year": w_warp, "events_per_year": events,
                         "peak_source": f"{sm_count} SMs x 4 issue/clk x {sm_max:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json)",
                         "ncu_issue_active_pct": prof.get("issue_active_pct"),
                         "traffic_source": prof.get("source"),
                         "hbm": {"achieved": (prof.get("dram_bytes_per_launch") or 0.0) / k_s / 1e9, "peak": peaks.get("hbm_gbs"),
                                 "unit": "GB/s", "frac": (prof.get("dram_bytes_per_launch") or 0.0) / k_s / 1e9 / float(peaks.get("hbm_gbs") or 6547.5)},
                         "hbm_note": "HBM traffic is per-launch accumulators only; not the bound (SURVEY 8d)"},
            "results": {"years": idx.years, "lole_h_per_yr": idx.lole, "lole_se": idx.lole_se,
                        "eens_mwh_per_yr": idx.eens, "eens_se": idx.eens_se, "lolf_occ_per_yr": idx.lolf,
                        "lold_h": idx.lold, "analytical_lole": 9.3677375218, "analytical_eens": 1176.181257},
            "wall_ms_per_step": wall_ms / args.steps,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rate, yrs, secs = cpu_reference_rate(12.0, threads)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{yrs} RTS-79 system-years ({yrs // threads} per thread), literal "
                                              f"hour/unit loop of PSA.jl:214-269, {secs:.1f} s"}
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
