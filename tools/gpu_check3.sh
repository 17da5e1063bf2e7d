mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 300 python -m pytest tests/test_gpu_zipenh.py tests/test_gpu_graphs.py -m gpu -x -q > $O/pytest_zip.log 2>&1; echo "pytest zipenh rc=$?"; tail -3 $O/pytest_zip.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_sa.json 2> $O/bench.err
tail -1 $O/bench_sa.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms_per_step']
print(round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {n:round(v,2) for n,v in k.items() if v>0.2})"
