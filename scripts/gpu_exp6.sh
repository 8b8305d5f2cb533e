echo "=== tests"; timeout 600 python -m pytest tests/test_gpu_lstm.py -q -m gpu --timeout 150 -x 2>&1 | tail -3
for d in 0 1 2 3; do echo "=== dbg=$d"; GR_TC_DBG=$d timeout 120 python scripts/trace_tc.py 2>&1 | grep -v "globaltimer\|slowest\|poll_done  "; done
echo "=== perf"; timeout 300 python scripts/lstm_perf.py
