mkdir -p gpurun_out
# (1) launch list of the bench command itself (first 1200 launches: warm-up steps of the depth-24, B=64 workload)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_ss.csv \
    python bench.py --model mf2ss --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
# (2) full metric set for every kernel of one layer at B=64 (second repetition: warm); raw page exported here,
#     the report itself (tens of MB) stays on the box
timeout 600 ncu --set full --clock-control none --launch-skip 31 -c 31 -f -o /tmp/ss_full \
    python tools/prof_ss.py 1 64 2 > gpurun_out/prof_ss.log 2>&1
echo "full rc=$?"
ncu -i /tmp/ss_full.ncu-rep --page raw --csv > gpurun_out/ss_full_raw.csv 2>/dev/null
ncu -i /tmp/ss_full.ncu-rep --page details --csv > gpurun_out/ss_full_details.csv 2>/dev/null
ls -la gpurun_out/ | head
