# MossFormerGAN-SE-16K evidence run (one GPU): parity tests, bench line, ncu launch list, full metric set for one block, mixed stream.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_mfgan.py -x -q -s > gpurun_out/mfgan_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/mfgan_tests.log
timeout 240 python bench.py --model mfgan --batch 32 --steps 3 --warmup 3 > gpurun_out/bench_mfgan.json 2> gpurun_out/bench_mfgan.err; echo "bench rc=$?"
# (1) launch list: the first run (512 launches) of the 6-block model on 8 windows
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 560 --csv --log-file gpurun_out/launches_gan.csv \
    python tools/time_mfgan.py 8 6 > gpurun_out/time_gan_under_ncu.log 2>&1
echo "launch list rc=$?"
# (2) full metric set for one block (intra path, inter path, triple attention) at 2 windows, second run (warm).
#     NOTE: the round-1 run of this step at 8 windows x 67 launches hit its 400 s limit (each launch is replayed ~40 times and the
#     attention GEMMs are 0.5-0.9 ms each) and spent the rest of the round's GPU minutes: keep it to 2 windows and 40 launches.
timeout 150 ncu --set full --clock-control none --launch-skip 548 -c 40 -f -o /tmp/gan_full \
    python tools/time_mfgan.py 2 6 > gpurun_out/prof_gan.log 2>&1
echo "full rc=$?"
ncu -i /tmp/gan_full.ncu-rep --page raw --csv > gpurun_out/gan_full_raw.csv 2>/dev/null
timeout 200 python tools/bench_mixed.py --chunks 128 --steps 2 --warmup 1 > gpurun_out/bench_mixed.json 2> gpurun_out/bench_mixed.err; echo "mixed rc=$?"
tail -c 1500 gpurun_out/bench_mfgan.json; echo; tail -c 600 gpurun_out/bench_mixed.json; tail -3 gpurun_out/bench_mixed.err
