# full-set ncu capture (SASS source counters) of the first ff_in (N = 256, K = 64) and ff_out (N = 64, K = 256) GEMMs of a ZipEnhancer pass
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel --launch-skip 6 --launch-count 1 -f -o /tmp/ffi python tools/run_once.py --model zipenh --batch 64 --runs 1 > $O/ncu_ffi.log 2>&1
ncu -i /tmp/ffi.ncu-rep --page raw --csv > $O/ffi_raw.csv 2>/dev/null
ncu -i /tmp/ffi.ncu-rep --page source --csv --print-source sass > $O/ffi_src.csv 2>/dev/null
ncu -i /tmp/ffi.ncu-rep --page details > $O/ffi_details.txt 2>/dev/null
ls -la $O/ffi*; tail -3 $O/ncu_ffi.log
