mkdir -p gpurun_out
timeout 200 python scripts/trace_a32.py dw 2>&1 | tee gpurun_out/r2_a32_trace_dw.log | tail -22
GR_A32_NVG=1 timeout 200 python scripts/trace_a32.py dw 2>&1 | tee gpurun_out/r2_a32_trace_dw_nvg1.log | tail -22
