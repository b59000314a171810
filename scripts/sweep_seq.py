"""Launch-geometry sweep of the sequential kernel (RTS-79): years/s vs segment length and warps/block."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from powersystemsreliabilityassessment_b200 import Engine, rts79

years = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
cap, mttf, mttr = rts79.units()
load = rts79.load_curve_int()
SEGS = [int(x) for x in os.environ.get('SEGS', '2944,2208,1760,1472,1120').split(',')]
WPBS = [int(x) for x in os.environ.get('WPBS', '16,24').split(',')]
rows = []
for seg in SEGS:
    for wpb in WPBS:
        try:
            with Engine(seg_hours=seg, warps_per_block=wpb) as e:
                e.set_system(cap, mttf, mttr); e.set_load(load)
                e.seq_mc(200_000, seed=1)
                best = min(e.seq_mc(years, seed=2 + i).kernel_ms for i in range(3))
                r = e.seq_mc(years, seed=2)
                c = e.last_counters(); rows.append(dict(seg=seg, wpb=wpb, ms=round(best, 2), Myps=round(years / best / 1e3, 1), waves=c['waves'] / years, jobs=c['jobs'] / years, ahead=c['ahead_jobs'] / years, runs=c['resolved_runs'] / years, pend_max=c['pend_max']))
                print(rows[-1], flush=True)
        except Exception as ex:
            print("fail", seg, wpb, ex, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/sweep_seq.json", "w"), indent=1)
