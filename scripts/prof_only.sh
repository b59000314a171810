set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:seq_ -s 1 -c 1 -f -o gpurun_out/prof_fast python scripts/profile_seq.py 1e6 > gpurun_out/prof_fast.log 2>&1; tail -3 gpurun_out/prof_fast.log
