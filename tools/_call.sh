mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_zhgtcrn.py -m gpu -q -s > gpurun_out/c2_hg.log 2>&1; echo "pytest rc=$?"
grep -v "^$" gpurun_out/c2_hg.log | tail -40
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c2_bench_zip.log 2>&1
python - <<PY
import json
l=open("gpurun_out/c2_bench_zip.log").read().strip().splitlines()[-1]
try:
    d=json.loads(l); print(d["value"], d["ms_per_step"], d["e2e"]["value"]); print(d["kernels_ms_per_step"])
except Exception as e: print(l[-2000:])
PY
