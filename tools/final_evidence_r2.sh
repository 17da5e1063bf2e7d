# Round-2 end evidence on one B200: the default bench exactly as the driver runs it (+ the reference arm), one bench line per
# workload, the ncu launch list of the default bench command, a full-set ncu capture of one Zipformer layer + one dense conv,
# launch lists of the families added this round.  Nothing printed under ncu is a bench value.
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
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
# launch list of the default bench command (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/zipenh_b64_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
# full metric set: one dense conv + one full-resolution Zipformer layer of the warm second pass (launch indices of zip::forward)
timeout 900 ncu --set full --clock-control none --launch-skip 290 -c 60 -f -o /tmp/zip_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 64 > $O/zip_full.log 2>&1
ncu -i /tmp/zip_full.ncu-rep --page raw --csv > $O/zipenh_b64_ncu_raw.csv 2>/dev/null
for m in hgtcrn dfsmn ulunas mfgan; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/${m}_launches.csv python bench.py --model $m --steps 1 --warmup 1 --no-cpu-baseline > $O/${m}_under_ncu.log 2>&1
done
timeout 600 ncu --set full --clock-control none -k regex:"wpe_kernel|iva_kernel" -c 2 -f -o /tmp/hg_full python bench.py --model hgtcrn --steps 1 --warmup 1 --no-cpu-baseline > $O/hg_full.log 2>&1
ncu -i /tmp/hg_full.ncu-rep --page raw --csv > $O/hgtcrn_ncu_raw.csv 2>/dev/null
ls -la $O | head -40
