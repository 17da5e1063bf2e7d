mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_mf2se.py -m gpu -x -q -s > gpurun_out/mf2.log 2>&1; echo "pytest rc=$?"
grep -E "FAIL|max\|err\||LSB|passed|failed|Error" gpurun_out/mf2.log | head -20
timeout 300 python bench.py --model mf2se --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mf2.log 2>&1
python - <<PY
import json
l=open("gpurun_out/bench_mf2.log").read().strip().splitlines()[-1]
try:
    d=json.loads(l); print(d["value"], d["ms_per_step"], d["e2e"]["value"]); print(d["kernels_ms_per_step"]); print(d["roofline"])
except Exception as e: print(l[-2000:])
PY
