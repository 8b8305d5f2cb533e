mkdir -p gpurun_out
timeout 600 python scripts/step_probe.py 2>&1 | tee gpurun_out/r2_step_probe0.log | tail -12
