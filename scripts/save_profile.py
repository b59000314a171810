"""Summarise the last gpurun visit into profiles/ (tracked): bench line, launch list, ncu details + SASS breakdown."""
import csv, json, os, shutil, subprocess, sys
tag = sys.argv[1]
rep = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/prof_seq.ncu-rep"
years = float(sys.argv[3]) if len(sys.argv) > 3 else 1e6
os.makedirs("profiles", exist_ok=True)
for src, dst in (("gpurun_out/bench.json", f"profiles/{tag}_bench.json"), ("gpurun_out/bench_ref.json", f"profiles/{tag}_bench_reference.json"),
                 ("gpurun_out/launches.csv", f"profiles/{tag}_launches_bench.csv")):
    if os.path.exists(src):
        shutil.copy(src, dst)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["Kernel Name", "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg.per_second",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active"]
summ = {}
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        summ[k] = {"unit": units[i], "values": [r[i] for r in data]}
def num(k, j=0):
    return float(summ[k]["values"][j].replace(",", ""))
to_b = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
dram = num("dram__bytes_read.sum") * to_b[summ["dram__bytes_read.sum"]["unit"]] + num("dram__bytes_write.sum") * to_b[summ["dram__bytes_write.sum"]["unit"]]
roof = {"source": f"{tag}: ncu --set full, {rep}, {years:.0f} RTS-79 system-years per launch",
        "dram_bytes_per_launch": dram, "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warp_inst_per_year": num("smsp__inst_executed.sum") / years, "kernel_ms": num("gpu__time_duration.sum"),
        "avg_active_threads_per_inst": num("smsp__thread_inst_executed_per_inst_executed.ratio")}
json.dump({"roofline": roof, "metrics": summ}, open(f"profiles/{tag}_seq_kernel_ncu_summary.json", "w"), indent=1)
json.dump(roof, open("profiles/roofline_latest.json", "w"), indent=1)
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
open("/tmp/_sass.csv", "w").write(sass)
out = subprocess.run([sys.executable, "scripts/sass_breakdown.py", "/tmp/_sass.csv", str(years)], capture_output=True, text=True).stdout
open(f"profiles/{tag}_seq_kernel_sass_breakdown.txt", "w").write(out)
print(json.dumps(roof, indent=1))
