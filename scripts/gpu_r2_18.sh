mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x 2>&1 | tail -4
timeout 600 python scripts/step_probe.py 2>&1 | tee gpurun_out/r2_step_probe1.log | tail -12
