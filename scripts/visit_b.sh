# GPU visit: full parity tests, timing of the sequential kernels, cheap ncu counters
set -x
mkdir -p gpurun_out
make -C oracle -s
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python scripts/profile_seq.py 1e7 2>&1 | tail -3
timeout 300 python scripts/profile_wide.py 2e5 2>&1 | tail -2
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,launch__registers_per_thread,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -k regex:seq_fast -s 1 -c 1 python scripts/profile_seq.py 1e6 2>&1 | grep -E "inst_executed|issue_active|duration|registers"
timeout 600 ncu --metrics $M --clock-control none -k regex:seq_wide -s 1 -c 1 python scripts/profile_wide.py 1e5 2>&1 | grep -E "inst_executed|issue_active|duration|registers"
