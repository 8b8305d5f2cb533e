mkdir -p gpurun_out
timeout 200 python scripts/store_probe.py 2>&1 | tee gpurun_out/r2_store_probe2.log | tail -4
GR_A32_EPI=tma timeout 200 python scripts/store_probe.py 2>&1 | tail -3 | head -1
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_lstm.py tests/test_gpu_models.py -q -m gpu --timeout 300 2>&1 | tail -3
