mkdir -p gpurun_out
timeout 200 python scripts/trace_a32.py dw 2>&1 | grep -E "^dw ms|cycles per"
GR_A32_SPLITS=12 timeout 200 python scripts/trace_a32.py dw 2>&1 | grep -E "^dw ms"
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_lstm.py tests/test_gpu_models.py -q -m gpu --timeout 300 2>&1 | tail -3
timeout 300 python scripts/step_probe.py 2>&1 | tail -8
