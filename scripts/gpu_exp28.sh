timeout 250 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_models.py -q -m gpu --timeout 60 -x 2>&1 | tail -3
SHAPES=256x100,64x100,32x100 timeout 100 python scripts/lstm_perf.py
GR_TOWER_STREAMS=0 timeout 100 python scripts/step_breakdown.py 2>&1 | grep "lstm_recurrence"
