mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 300 python -m pytest tests/test_gpu_zipenh.py -m gpu -x -q > $O/pytest_zip.log 2>&1; echo "pytest zipenh rc=$?"; tail -3 $O/pytest_zip.log
for p in 0 4; do
ADN_TC_PROBE=$p timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_p$p.json 2> $O/bench_p$p.err
tail -1 $O/bench_p$p.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms_per_step']
print('probe $p', round(d['ms_per_step'],2), {n:round(v,2) for n,v in k.items() if v>0.9})"
done
