# gpu suite + the default bench (per-kernel table)
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
tail -1 $O/bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms_per_step']
print(round(d['ms_per_step'],2), {n:round(v,2) for n,v in k.items() if v>0.9})"
