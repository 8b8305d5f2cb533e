mkdir -p gpurun_out
timeout 200 python scripts/store_probe.py 2>&1 | tee gpurun_out/r2_store_probe.log | tail
