mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
echo "=== tests default (FAST)"; timeout 400 python -m pytest tests/test_gpu_lstm.py -q -m gpu --timeout 150 2>&1 | tail -5
echo "=== tests PRECISE"; GR_LSTM_TC_PRECISE=1 timeout 400 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_models.py -q -m gpu --timeout 150 2>&1 | tail -5
echo "=== perf"; timeout 300 python scripts/lstm_perf.py
echo "=== trace FAST"; timeout 120 python scripts/trace_tc.py
echo "=== trace PRECISE"; GR_LSTM_TC_PRECISE=1 timeout 120 python scripts/trace_tc.py
