mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lstm.py -x -q --timeout 300 2>&1 | tail -3
(for sy in flags ctr; do echo "== GR_TCU_SYNC=$sy"; GR_TCU_SYNC=$sy SHAPES=256x500,128x500,64x500,32x500,256x300 timeout 300 python scripts/trace_tcu.py | grep -E "^B|period|tma_poll_done|mma_first_full|epi_red"; done) 2>&1 | tee gpurun_out/r2_tcu_flags_trace.log
for sy in flags ctr; do echo "== GR_TCU_SYNC=$sy"; GR_TCU_SYNC=$sy timeout 300 python scripts/lstm_perf.py 2>&1 | tail -12; done | tee gpurun_out/r2_tcu_flags_time.log
