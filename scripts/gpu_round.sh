# one GPU visit: tests, smoke, bench (both arms), ncu launch list + full capture of the top kernel
set -x
mkdir -p gpurun_out
make -C oracle -s
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 --ref-budget 30 > gpurun_out/bench_ref.json 2>&1; tail -c 1500 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 python scripts/configs_report.py > gpurun_out/configs.log 2>&1; tail -3 gpurun_out/configs.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:seq_fast_kernel -s 1 -c 2 -f -o gpurun_out/prof_seq python scripts/profile_seq.py 1e6 > gpurun_out/prof_seq.log 2>&1; tail -3 gpurun_out/prof_seq.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:seq_wide_kernel -s 1 -c 1 -f -o gpurun_out/prof_wide python scripts/profile_wide.py 1e5 > gpurun_out/prof_wide.log 2>&1; tail -3 gpurun_out/prof_wide.log
