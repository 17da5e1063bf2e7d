mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 400 python -m pytest tests -m gpu -x -q -s -k "full_depth_reference" > $O/pytest_fd.log 2>&1; echo "pytest rc=$?"; grep -E "full depth|depth 6|passed|failed|Error" $O/pytest_fd.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
