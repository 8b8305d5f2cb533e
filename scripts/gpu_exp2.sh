echo "=== trace PRECISE keep=0"; GR_LSTM_TC_PRECISE=1 timeout 120 python scripts/trace_tc.py
echo "=== trace PRECISE keep=1"; KEEP=1 GR_LSTM_TC_PRECISE=1 timeout 120 python scripts/trace_tc.py
