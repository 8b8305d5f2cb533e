mkdir -p gpurun_out
timeout 200 python scripts/tcu_check.py check 2>&1 | tail -3
SHAPES=256x500 timeout 200 python scripts/trace_tcu.py 2>&1 | tee gpurun_out/r2_tcu_trace2.log | tail -40
timeout 300 python scripts/tcu_check.py time 2>&1 | grep tcu | tee gpurun_out/r2_tcu_time2.log
