# quick GPU visit for the sequential kernel: parity tests, timing, and a cheap ncu instruction count
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_seq.py -x -q 2>&1 | tail -5
timeout 300 python scripts/profile_seq.py 1e7 2>&1 | tail -4
timeout 600 ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,launch__registers_per_thread --clock-control none -k regex:seq_fast -s 1 -c 1 python scripts/profile_seq.py 1e6 2>&1 | grep -E "inst_executed|issue_active|duration|registers" 
