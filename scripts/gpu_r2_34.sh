mkdir -p gpurun_out
(for off in 0 1000 2000; do
echo "== ffwd dropout nvg=4 off=$off"; GR_A32_TRACE_OFF=$off TRACE_MASK_SCALE=2 timeout 200 python scripts/trace_a32.py ffwd
done
echo "== ffwd generic nvg=1 off=200"; GR_A32_TRACE_OFF=200 timeout 200 python scripts/trace_a32.py ffwd
echo "== dw dropout nvg=4 off=200"; GR_A32_TRACE_OFF=200 TRACE_MASK_SCALE=2 timeout 200 python scripts/trace_a32.py dw
echo "== fwd K=1000 off=100"; GR_A32_TRACE_OFF=100 timeout 200 python scripts/trace_a32.py fwd
) 2>&1 | grep -v "^  -\|half\|fence done\|epi:" | tee gpurun_out/r2_binmask_trace2.log
