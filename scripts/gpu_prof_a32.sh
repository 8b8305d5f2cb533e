mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_a32 -s 1 -c 1 -f -o gpurun_out/prof_a32 python scripts/micro.py a32 > gpurun_out/ncu_a32.log 2>&1; echo "exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_a32 -s 1 -c 1 -f -o gpurun_out/prof_a32t python scripts/micro.py a32t > gpurun_out/ncu_a32t.log 2>&1; echo "exit $?"
tail -3 gpurun_out/ncu_a32.log
