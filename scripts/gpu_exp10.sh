timeout 200 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_lstm.py -q -m gpu --timeout 60 -x 2>&1 | tail -5
timeout 100 python scripts/trace_a32.py dw
timeout 100 python scripts/trace_a32.py fwd
