GR_A32_EPI=wide timeout 300 python scripts/store_probe2.py 2>&1 | grep "F=40" | grep -v tma
timeout 300 python scripts/store_probe2.py 2>&1 | grep "F=40" | grep -v tma
GR_A32_EPI=wide timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_lstm.py -q -m gpu --timeout 300 2>&1 | tail -3
