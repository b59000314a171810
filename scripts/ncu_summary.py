"""Print (and optionally save as JSON) the counters of one `ncu --set full` capture that the roofline discussion uses.
usage: python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep [profiles/r02x_kernel_ncu_summary.json]"""
import csv, json, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, val = rows[0], rows[1], rows[-1]
want = ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_", "sm__pipe_", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "l1tex__data_pipe_lsu_wavefronts_mem_shared",
        "smsp__inst_executed_op_shared", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__average_warp", "smsp__average_warps_issue_stalled", "Kernel Name",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__throughput.avg.pct", "l1tex__throughput.avg.pct", "smsp__thread_inst_executed_per_inst_executed")
out = {}
for h, u, v in zip(hdr, units, val):
    if any(w in h for w in want) and "_pred_on" not in h:
        out[h] = v if not u else f"{v} {u}"
for k in sorted(out): print(f"{k:110s} {out[k]}")
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1, sort_keys=True)
