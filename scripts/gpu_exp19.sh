timeout 200 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_lstm.py tests/test_gpu_models.py -q -m gpu --timeout 60 -x 2>&1 | tail -5
timeout 100 python scripts/trace_a32.py dw 2>&1 | grep "ms\|cycles per"
GR_A32_NVG=1 timeout 100 python scripts/trace_a32.py dw 2>&1 | grep "ms\|cycles per"
GR_TOWER_STREAMS=0 timeout 100 python scripts/step_breakdown.py 2>&1 | tail -24
