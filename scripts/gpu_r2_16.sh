mkdir -p gpurun_out
./scripts/micro/ffma_rate 2>&1 | tee gpurun_out/r2_micro_ffma_rate.log
