echo "=== tests"; timeout 600 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_models.py -q -m gpu --timeout 150 -x 2>&1 | tail -15
echo "=== perf"; timeout 300 python scripts/lstm_perf.py
echo "=== trace keep=0"; timeout 120 python scripts/trace_tc.py
echo "=== trace keep=1"; KEEP=1 timeout 120 python scripts/trace_tc.py
