mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
timeout 200 compute-sanitizer --tool racecheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > $O/racecheck_smoke.txt 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|hazard" $O/racecheck_smoke.txt | head -6
timeout 250 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_mfgan.py -m gpu -x -q -k "fixture" > $O/sanitizer_mfgan.txt 2>&1; echo "memcheck gan rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid" $O/sanitizer_mfgan.txt | head -6
