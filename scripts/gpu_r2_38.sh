mkdir -p gpurun_out
(SHAPES=256x500,32x500,256x300 timeout 300 python scripts/trace_tcu.py
for d in 2 4 6; do echo "== GR_TCU_DBG=$d"; GR_TCU_DBG=$d SHAPES=256x500,32x500 timeout 300 python scripts/trace_tcu.py | grep -E "^B|period"; done) 2>&1 | tee gpurun_out/r2_tcu_trace5.log
