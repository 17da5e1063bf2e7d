mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_zdfsmn.py tests/test_gpu_mfgan.py -m gpu -x -q -s > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/c1_pytest.log
timeout 200 python tools/time_mfgan.py 16 6 > gpurun_out/c1_mfgan_time.log 2>&1; head -30 gpurun_out/c1_mfgan_time.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench_zip.log 2>&1
python - <<PY
import json
l=open("gpurun_out/c1_bench_zip.log").read().strip().splitlines()[-1]
try:
    d=json.loads(l); print(d["value"], d["ms_per_step"], d["e2e"]["value"]); print(d["kernels_ms_per_step"]); print(d["roofline"])
except Exception as e: print(l[-2000:])
PY
timeout 100 python bench.py --model dfsmn --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-600
