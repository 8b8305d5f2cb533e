mkdir -p gpurun_out
timeout 300 python scripts/small_perf.py time 2>&1 | tee gpurun_out/r2_small_time0.log | tail -12
timeout 300 python scripts/small_perf.py check 2>&1 | tee gpurun_out/r2_small_check0.log | tail -12
