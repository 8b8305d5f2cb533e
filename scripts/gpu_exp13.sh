timeout 200 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_models.py -q -m gpu --timeout 60 -x 2>&1 | tail -5
SHAPES=256x100,32x100 timeout 100 python scripts/lstm_perf.py
GR_LSTM_IMPL=tc SHAPES=256x100,32x100 timeout 100 python scripts/lstm_perf.py
