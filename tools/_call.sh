mkdir -p gpurun_out
timeout 200 python tools/time_mfgan.py 16 6 > gpurun_out/c9_mfgan_time.log 2>&1; head -12 gpurun_out/c9_mfgan_time.log
timeout 300 python -m pytest tests/test_gpu_mfgan.py tests/test_gpu_zdfsmn.py tests/test_gpu_zmfgan_resample.py -m gpu -x -q 2>&1 | tail -2
