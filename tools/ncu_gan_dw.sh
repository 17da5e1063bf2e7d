mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dw_conv_seq_kernel --launch-skip 8 --launch-count 4 -f -o /tmp/gdw python tools/run_once.py --model mfgan --batch 16 --runs 1 > $O/ncu_gdw.log 2>&1
ncu -i /tmp/gdw.ncu-rep --page raw --csv > $O/gdw_raw.csv 2>/dev/null
ncu -i /tmp/gdw.ncu-rep --page source --csv --print-source sass > $O/gdw_src.csv 2>/dev/null
ncu -i /tmp/gdw.ncu-rep --page details > $O/gdw_details.txt 2>/dev/null
ls -la $O/gdw*; tail -3 $O/ncu_gdw.log
