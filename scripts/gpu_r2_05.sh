mkdir -p gpurun_out
SHAPES=${SHAPES:-256x500} timeout 200 python scripts/trace_tcu.py 2>&1 | tee gpurun_out/r2_tcu_trace3.log | tail -50
