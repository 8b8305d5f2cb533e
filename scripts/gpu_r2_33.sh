mkdir -p gpurun_out
(echo "== ffwd generic nvg=1"; timeout 200 python scripts/trace_a32.py ffwd
echo "== ffwd generic nvg=4"; GR_A32_NVG=4 timeout 200 python scripts/trace_a32.py ffwd
echo "== ffwd dropout nvg=4"; TRACE_MASK_SCALE=2 timeout 200 python scripts/trace_a32.py ffwd
echo "== dw generic nvg=4"; timeout 200 python scripts/trace_a32.py dw
echo "== dw dropout nvg=4"; TRACE_MASK_SCALE=2 timeout 200 python scripts/trace_a32.py dw) 2>&1 | grep -v "^  -\|half\|fence done\|epi:" | tee gpurun_out/r2_binmask_trace.log
