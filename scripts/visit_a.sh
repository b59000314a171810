# GPU visit: full parity tests, timing of the sequential kernels, instruction counts, one source-level capture
set -x
mkdir -p gpurun_out
make -C oracle -s
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python scripts/profile_seq.py 1e7 2>&1 | tail -3
timeout 300 python scripts/profile_wide.py 2e5 2>&1 | tail -2
timeout 600 ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,launch__registers_per_thread --clock-control none -k regex:seq_wide -s 1 -c 1 python scripts/profile_wide.py 1e5 2>&1 | grep -E "inst_executed|issue_active|duration|registers"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:seq_fast_kernel -s 1 -c 1 -f -o gpurun_out/prof_seq python scripts/profile_seq.py 1e6 > gpurun_out/prof_seq.log 2>&1; tail -3 gpurun_out/prof_seq.log
