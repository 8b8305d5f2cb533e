mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_ctc.py tests/test_golden.py tests/test_gpu_models.py -q -m gpu --timeout 150 2>&1 | tail -8
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ctc_loss_grad -s 2 -c 1 -f -o gpurun_out/prof_ctc4 python scripts/micro.py ctc > gpurun_out/ncu_ctc4.log 2>&1
echo "ctc prof exit $?"; ls -la gpurun_out/prof_ctc4.ncu-rep
