timeout 300 python scripts/store_probe2.py 2>&1 | tee gpurun_out/r2_store_probe3.log | tail -26
