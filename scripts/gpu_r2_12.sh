mkdir -p gpurun_out
timeout 300 python scripts/tcu_bwd_check.py check 2>&1 | tee gpurun_out/r2_tcu_bwd_check.log | tail -24
echo "exit $?"
timeout 300 python scripts/tcu_bwd_check.py time 2>&1 | tee gpurun_out/r2_tcu_bwd_time.log | tail -12
