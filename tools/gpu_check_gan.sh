mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 400 python -m pytest tests/test_gpu_mfgan.py tests/test_gpu_zmfgan_resample.py tests/test_gpu_zdfsmn.py tests/test_gpu_zulunas.py -m gpu -x -q > $O/pytest_gan.log 2>&1; echo "pytest gan rc=$?"; tail -3 $O/pytest_gan.log
for v in seq; do
ADN_GAN_DW=$v timeout 300 python bench.py --model mfgan --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_gan_$v.json 2> $O/bench_gan_$v.err
tail -1 $O/bench_gan_$v.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms_per_step']
print('$v', round(d['ms_per_step'],2), {n:round(v,2) for n,v in k.items() if v>10})"
done
