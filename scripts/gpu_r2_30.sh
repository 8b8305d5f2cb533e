timeout 300 python scripts/store_probe3.py 2>&1 | tee gpurun_out/r2_store_probe5.log | tail -20
