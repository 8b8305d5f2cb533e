# binary-mask producers of gemm_a32: parity + timing of the fusion projection / dW with and without them
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_abi.py -x -q --timeout 300 2>&1 | tail -5
timeout 300 python scripts/binmask_probe.py 2>&1 | tee gpurun_out/r2_binmask_probe.log
