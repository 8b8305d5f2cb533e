mkdir -p gpurun_out
timeout 300 python scripts/tcu_check.py check 2>&1 | tee gpurun_out/r2_tcu_check.log | tail -40
echo "exit $?"
timeout 300 python scripts/tcu_check.py time 2>&1 | tee gpurun_out/r2_tcu_time.log | tail -40
