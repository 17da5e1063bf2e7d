# Round-end evidence: one bench line per workload (not under a profiler) + compute-sanitizer memcheck on small runs
mkdir -p gpurun_out/final
timeout 400 python bench.py > gpurun_out/final/bench_gtcrn.json 2>gpurun_out/final/err_gtcrn.log
timeout 400 python bench.py --model mbr --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/final/bench_mbr.json 2>gpurun_out/final/err_mbr.log
timeout 400 python bench.py --model mf2se --steps 10 --warmup 3 > gpurun_out/final/bench_mf2se.json 2>gpurun_out/final/err_mf2se.log
timeout 400 python bench.py --model mf2se --matmul bf16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/final/bench_mf2se_bf16.json 2>gpurun_out/final/err_mf2se_bf16.log
timeout 600 python bench.py --model mf2ss --steps 5 --warmup 3 --cpu-baseline-chunks 3 > gpurun_out/final/bench_mf2ss.json 2>gpurun_out/final/err_mf2ss.log
for f in gpurun_out/final/bench_*.json; do tail -1 $f | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['model'], d['dtype'], round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['kernel'], round(d['roofline']['frac'],3))"; done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/prof_mf2.py 2 2 1 bf16 > gpurun_out/final/sanitizer_mf2se_bf16.log 2>&1; echo "memcheck mf2se bf16 rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/sanitizer_gtcrn_smoke.log 2>&1; echo "memcheck gtcrn smoke rc=$?"
grep -h "ERROR SUMMARY" gpurun_out/final/sanitizer_*.log
