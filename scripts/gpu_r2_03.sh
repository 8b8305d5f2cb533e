mkdir -p gpurun_out
timeout 200 python scripts/trace_tcu.py 2>&1 | tee gpurun_out/r2_tcu_trace.log | tail -80
