mkdir -p gpurun_out
timeout 300 python scripts/small_perf.py check 2>&1 | tee gpurun_out/r2_small_check1.log | tail -12
timeout 300 python scripts/small_perf.py time 2>&1 | tee gpurun_out/r2_small_time1.log | tail -12
GR_SMALL_BS=2 BS=32,64,256 timeout 300 python scripts/small_perf.py time 2>&1 | tee -a gpurun_out/r2_small_time1.log | tail -4
GR_SMALL_BS=4 BS=32,64,128 timeout 300 python scripts/small_perf.py time 2>&1 | tee -a gpurun_out/r2_small_time1.log | tail -4
