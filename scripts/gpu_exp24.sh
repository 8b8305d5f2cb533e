timeout 250 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_lstm.py tests/test_gpu_models.py -q -m gpu --timeout 60 -x 2>&1 | tail -3
for eb in 6 4 2; do echo "epi bufs $eb"; GR_A32_EPI_BUFS=$eb GR_TOWER_STREAMS=0 timeout 100 python scripts/step_breakdown.py 2>&1 | grep "TFLOP/s alg" | head -3; done
