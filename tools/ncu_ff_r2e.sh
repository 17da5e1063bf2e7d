# full-set ncu capture (SASS source counters) of the first ff_in (N = 256, K = 64) and ff_out (N = 64, K = 256) GEMMs of a ZipEnhancer pass
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel --launch-skip 7 --launch-count 1 -f -o /tmp/ffo python tools/run_once.py --model zipenh --batch 64 --runs 1 > $O/ncu_ffo.log 2>&1
ncu -i /tmp/ffo.ncu-rep --page raw --csv > $O/ffo_raw.csv 2>/dev/null
ncu -i /tmp/ffo.ncu-rep --page source --csv --print-source sass > $O/ffo_src.csv 2>/dev/null
ncu -i /tmp/ffo.ncu-rep --page details > $O/ffo_details.txt 2>/dev/null
ls -la $O/ffo*; tail -3 $O/ncu_ffo.log
