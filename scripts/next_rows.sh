# measurement of the "next" rows: detailed MC and multi-area kernels -- throughput + issue counters
set -x
mkdir -p gpurun_out
timeout 600 python scripts/configs_report.py > gpurun_out/configs.log 2>&1; tail -40 gpurun_out/configs.log | grep -A6 "detailed_mc_2e6\|multi_area"
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -k regex:"detailed|multi_area" -c 8 python scripts/configs_report.py 2>&1 | grep -E "^  [a-z_]+.*\(|inst_executed|issue_active|duration|registers|warps_active"
