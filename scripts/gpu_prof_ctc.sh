mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ctc_loss_grad -s 2 -c 1 -f -o gpurun_out/prof_ctc2 python scripts/micro.py ctc > gpurun_out/ncu_ctc2.log 2>&1
echo "ctc prof exit $?"; ls -la gpurun_out/prof_ctc2.ncu-rep
