mkdir -p gpurun_out
for d in 0 1 2 4 6; do
echo "== GR_TCU_DBG=$d"
GR_TCU_DBG=$d SHAPES=256x500 timeout 200 python scripts/trace_tcu.py 2>&1 | grep -E "period" 
done | tee gpurun_out/r2_tcu_dbg.log
