mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
timeout 200 ncu --set full --import-source on --clock-control none -k regex:attn_apply_mma_kernel --launch-skip 3 --launch-count 1 -f -o /tmp/sa python tools/run_once.py --model zipenh --batch 64 --runs 1 > $O/ncu_sa.log 2>&1
ncu -i /tmp/sa.ncu-rep --page source --csv --print-source sass > $O/sa_src.csv 2>/dev/null
ncu -i /tmp/sa.ncu-rep --page details > $O/sa_details.txt 2>/dev/null
tail -2 $O/ncu_sa.log
