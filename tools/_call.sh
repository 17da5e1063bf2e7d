mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_zhgtcrn.py tests/test_gpu_zdfsmn.py -m gpu -q -s 2>&1 | grep -E "window|hgtcrn|passed|failed|Error|error" | head -12
for mdl in hgtcrn dfsmn; do
timeout 200 python bench.py --model $mdl --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$mdl', d['config']['batch_per_gpu'], round(d['ms_per_step'],2), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), dict(list(d['kernels_ms_per_step'].items())[:8]))"
done
