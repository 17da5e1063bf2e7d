# GEMM bottleneck probe on the default workload (timings only: probe runs produce wrong numbers by design) + the gpu suite
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
for p in 0 1 2 4 7; do
  ADN_TC_PROBE=$p timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/probe_$p.json 2> $O/probe_$p.err
  tail -1 $O/probe_$p.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms_per_step']
print('probe $p', round(d['ms_per_step'],2), {n:round(v,2) for n,v in k.items() if v>1.0})"
done
