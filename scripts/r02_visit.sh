# round-2 GPU visit: tests, timings of the sequential kernels (variants), optional bench / ncu.  usage: bash scripts/r02_visit.sh [tests] [time] [bench] [ncuwide] [ncufast] [ref]
set -x
mkdir -p gpurun_out
make -C oracle -s
for what in "$@"; do
case $what in
tests)  timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ;;
tests_all)  timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 ;;
smoke)  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ;;
time)   timeout 600 python scripts/r02_time.py 2>&1 | tee gpurun_out/r02_time.log | tail -40 ;;
bench)  timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err ;;
ref)    timeout 300 python bench.py --impl reference --steps 3 --warmup 1 --ref-budget 30 > gpurun_out/bench_ref.json 2>&1; tail -c 1500 gpurun_out/bench_ref.json ;;
launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -3 gpurun_out/bench_under_ncu.log ;;
ncuwide) timeout 900 ncu --set full --clock-control none --import-source on -k regex:seq_wide_kernel -s 1 -c 1 -f -o gpurun_out/prof_wide python scripts/profile_wide.py 1e5 > gpurun_out/prof_wide.log 2>&1; tail -3 gpurun_out/prof_wide.log ;;
ncunonseq) timeout 900 ncu --set full --clock-control none --import-source on -k regex:nonseq_fast_kernel -s 1 -c 1 -f -o gpurun_out/prof_nonseq python scripts/profile_nonseq.py 1e8 > gpurun_out/prof_nonseq.log 2>&1; tail -3 gpurun_out/prof_nonseq.log ;;
ncufast) timeout 900 ncu --set full --clock-control none --import-source on -k regex:seq_fast_kernel -s 1 -c 1 -f -o gpurun_out/prof_seq python scripts/profile_seq.py 1e6 > gpurun_out/prof_seq.log 2>&1; tail -3 gpurun_out/prof_seq.log ;;
esac
done
