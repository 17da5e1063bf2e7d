# Round-2 end evidence on one B200 (after the TS-mode GEMM / elect.sync / epilogue work): the default bench exactly as the driver
# runs it (+ the reference arm), one bench line per workload, the ncu launch list of the default bench command, a full-set ncu
# capture of the dominant kernel (the K = 768 implicit-GEMM conv).  Nothing printed under ncu is a bench value.
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_default_reference_arm.json 2> $O/bench_ref.err
for m in gtcrn mf2se mbr mf2ss mfgan dfsmn ulunas hgtcrn; do
  timeout 600 python bench.py --model $m --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_$m.json 2> $O/err_$m.log
done
timeout 400 python bench.py --model mf2se --matmul bf16 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_mf2se_bf16.json 2> $O/err_mf2se_bf16.log
for f in $O/bench_*.json; do tail -1 $f | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d.get('impl','adn'), d['config'].get('model'), d['dtype'], round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), (d.get('roofline') or {}).get('kernel'), round((d.get('roofline') or {}).get('frac') or 0,3))
except Exception as e: print('bad', '$f', e)"; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/zipenh_b64_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel --launch-skip 0 --launch-count 4 -f -o /tmp/dense python tools/run_once.py --model zipenh --batch 64 --runs 1 > $O/ncu_dense.log 2>&1
ncu -i /tmp/dense.ncu-rep --page raw --csv > $O/zip_dense_ts_raw.csv 2>/dev/null
ncu -i /tmp/dense.ncu-rep --page details > $O/zip_dense_ts_details.txt 2>/dev/null
ls -la $O | head -40
