mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attention3_kernel --launch-skip 2 --launch-count 2 -o gpurun_out/c12_att3 python bench.py --model mbr --batch 8 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c12_ncu.log 2>&1
ncu -i gpurun_out/c12_att3.ncu-rep --page raw --csv > gpurun_out/c12_att3_raw.csv 2>/dev/null
ncu -i gpurun_out/c12_att3.ncu-rep --page source --csv --print-source sass > gpurun_out/c12_att3_src.csv 2>/dev/null
rm -f gpurun_out/c12_att3.ncu-rep
ls -la gpurun_out/c12*
