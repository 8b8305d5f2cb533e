mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ctc_loss_grad -s 2 -c 1 -f -o gpurun_out/r2_prof_ctc_v5 python scripts/micro.py ctc > gpurun_out/r2_ncu_ctc_v5.log 2>&1
echo "ctc prof exit $?"
