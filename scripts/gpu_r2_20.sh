timeout 500 python -m pytest tests/test_gpu_ctc.py tests/test_golden.py -q -m gpu --timeout 150 -x 2>&1 | tail -40
timeout 500 python -m pytest tests/test_gpu_ctc.py tests/test_golden.py -q -m gpu --timeout 150 2>&1 | grep -E "FAILED|passed|failed" | cut -c1-150
