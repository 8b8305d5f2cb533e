timeout 300 python scripts/store_probe2.py 2>&1 | tee gpurun_out/r2_store_probe4.log | grep "F=40"
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_lstm.py -q -m gpu --timeout 300 2>&1 | tail -3
BT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_a32_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_a32_l1 python scripts/store_probe.py > gpurun_out/r2_ncu_a32_l1.log 2>&1
echo "prof exit $?"
