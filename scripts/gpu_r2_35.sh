mkdir -p gpurun_out
(echo "== ffwd dropout nvg=4"; TRACE_MASK_SCALE=2 timeout 200 python scripts/trace_a32.py ffwd
echo "== ffwd generic nvg=1"; timeout 200 python scripts/trace_a32.py ffwd
echo "== dw dropout nvg=4"; TRACE_MASK_SCALE=2 timeout 200 python scripts/trace_a32.py dw
) 2>&1 | grep -v "^  -\|half\|fence done" | tee gpurun_out/r2_binmask_trace3.log
