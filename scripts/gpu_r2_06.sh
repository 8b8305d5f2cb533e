mkdir -p gpurun_out
timeout 200 python scripts/tcu_check.py check 2>&1 | tee gpurun_out/r2_tcu_check.log | grep -E 'FAIL|first bad|tcu check'
SHAPES=256x500,32x500 timeout 200 python scripts/trace_tcu.py 2>&1 | tee gpurun_out/r2_tcu_trace4.log | grep -E "tile|epi_|mma_|tma_|period"
timeout 300 python scripts/tcu_check.py time 2>&1 | grep tcu | tee gpurun_out/r2_tcu_time4.log
