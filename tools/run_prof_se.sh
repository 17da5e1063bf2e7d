mkdir -p gpurun_out
# (1) launch list of the default bench command (first 1000 launches: two warm-up steps of the depth-24, B=256 workload)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_se.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_se_under_ncu.log 2>&1
echo "launch list rc=$?"
# (2) full metric set for every kernel of one layer at B=256 (second repetition: warm)
timeout 600 ncu --set full --clock-control none --launch-skip 28 -c 28 -f -o /tmp/se_full \
    python tools/prof_mf2.py 1 256 2 > gpurun_out/prof_se.log 2>&1
echo "full rc=$?"
ncu -i /tmp/se_full.ncu-rep --page raw --csv > gpurun_out/se_full_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
