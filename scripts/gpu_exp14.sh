timeout 200 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_models.py -q -m gpu --timeout 60 -x 2>&1 | tail -5
timeout 100 python scripts/lstm_perf.py
timeout 60 python scripts/trace_tc.py 2>&1 | grep -v "globaltimer\|slowest\|poll_done  "
