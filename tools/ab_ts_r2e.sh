# TS (tensor-memory A operand) vs SS GEMM form: gpu suite, then the default bench in both forms
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu2.log
for ss in 0 1; do
  ADN_TC_SS=$ss timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/ts_ss$ss.json 2> $O/ts_ss$ss.err
  tail -1 $O/ts_ss$ss.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms_per_step']
print('ss=$ss', round(d['ms_per_step'],2), {n:round(v,2) for n,v in k.items() if v>1.0})"
done
