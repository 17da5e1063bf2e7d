# full-set ncu capture (with SASS source counters) of the ZipEnhancer dense-block implicit-GEMM convs (first GEMM launches of a pass:
# K = 384, 768, 1152, 1536, N = 64) and of the first ff_in / ff_out pair
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel --launch-skip 1 --launch-count 1 -f -o /tmp/dense python tools/run_once.py --model zipenh --batch 64 --runs 1 > $O/ncu_dense_ts.log 2>&1
ncu -i /tmp/dense.ncu-rep --page raw --csv > $O/dense_ts_raw.csv 2>/dev/null
ncu -i /tmp/dense.ncu-rep --page source --csv --print-source sass > $O/dense_ts_src.csv 2>/dev/null
ncu -i /tmp/dense.ncu-rep --page details > $O/dense_ts_details.txt 2>/dev/null
ls -la $O/dense*; tail -3 $O/ncu_dense_ts.log
