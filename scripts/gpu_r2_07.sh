mkdir -p gpurun_out
for d in 0 6; do
echo "== GR_TCU_DBG=$d"
GR_TCU_DBG=$d SHAPES=${SHAPES:-32x500} timeout 200 python scripts/trace_tcu.py 2>&1 | grep -E "period|mma_|tma_|epi_iter|epi_P" 
done | tee gpurun_out/r2_tcu_dbg32.log
