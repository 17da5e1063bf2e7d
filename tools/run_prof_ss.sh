mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mf2ss.py -m gpu -q -x -s -k "fold_window or shim" 2>&1 | grep -E "mf2ss L=|passed|failed|Error" 
# memcheck / racecheck on a small two-layer run (600 frames, 3 groups) through the C ABI
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/prof_ss.py 2 2 1 4808 > gpurun_out/sanitizer_memcheck_ss.log 2>&1; echo "memcheck rc=$?"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/prof_ss.py 1 1 1 4808 > gpurun_out/sanitizer_racecheck_ss.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/sanitizer_memcheck_ss.log; tail -3 gpurun_out/sanitizer_racecheck_ss.log
# (1) launch list of the bench command itself (first 1100 launches: two warm-up steps of the depth-24, B=64 workload)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_ss.csv \
    python bench.py --model mf2ss --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
# (2) full metric set for every kernel of one layer at B=64 (second repetition: warm); raw page exported here,
#     the report itself (tens of MB) stays on the box
timeout 600 ncu --set full --clock-control none --launch-skip 31 -c 31 -f -o /tmp/ss_full \
    python tools/prof_ss.py 1 64 2 > gpurun_out/prof_ss.log 2>&1
echo "full rc=$?"
ncu -i /tmp/ss_full.ncu-rep --page raw --csv > gpurun_out/ss_full_raw.csv 2>/dev/null
# (3) the bench line itself (not under a profiler)
timeout 600 python bench.py --model mf2ss --steps 5 --warmup 3 --cpu-baseline-chunks 3 > gpurun_out/bench_ss.log 2>&1
tail -1 gpurun_out/bench_ss.log | cut -c1-400
