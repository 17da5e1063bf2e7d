mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "tc vs|passed|failed|FAIL|Error|error|wave " gpurun_out/pytest_gpu.log | head -20
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
python - <<PY
import json
l=open("gpurun_out/bench.log").read().strip().splitlines()[-1]
try:
    d=json.loads(l); print(d["value"], d["ms_per_step"], d["e2e"]["value"]); print(d["kernels_ms_per_step"]); print(d["roofline"])
except Exception as e: print(l[-2000:])
PY
