timeout 250 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_lstm.py tests/test_gpu_models.py -q -m gpu --timeout 60 -x 2>&1 | tail -3
timeout 100 python scripts/trace_a32.py l1 2>&1 | grep "ms\|cycles per"
GR_A32_EPI=tma timeout 100 python scripts/trace_a32.py l1 2>&1 | grep "ms\|cycles per"
timeout 100 python scripts/trace_a32.py fwd 2>&1 | grep "ms\|cycles per"
GR_TOWER_STREAMS=0 timeout 100 python scripts/step_breakdown.py 2>&1 | grep "TFLOP/s alg\|sum of"
