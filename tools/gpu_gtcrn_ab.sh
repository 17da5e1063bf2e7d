mkdir -p gpurun_out/r2e
for g in 1 0 1 0; do
ADN_GRAPHS=$g timeout 200 python bench.py --model gtcrn --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('graphs=$g', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'], round(sum(d['kernels_ms_per_step'].values()),3))"
done
