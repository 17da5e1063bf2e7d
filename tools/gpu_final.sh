# last check of the round on one B200: the whole gpu suite, smoke(), the default bench and the GAN bench line
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err
tail -1 $O/bench_default.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms_per_step']
print('zipenh', round(d['ms_per_step'],2), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'], {n:round(v,2) for n,v in k.items() if v>0.4})"
timeout 300 python bench.py --model mfgan --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_mfgan.json 2> $O/bench_mfgan.err
tail -1 $O/bench_mfgan.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms_per_step']
print('mfgan', round(d['ms_per_step'],2), round(d['value'],1), {n:round(v,2) for n,v in k.items() if v>5})"
