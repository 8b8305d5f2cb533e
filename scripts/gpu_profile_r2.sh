# round-2 evidence: ncu launch list of the bench command + `--set full` captures of the two new recurrence kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 1 --warmup 1 --min-warmup 1 --skip-cpu --skip-ctc --skip-e2e > gpurun_out/r2_ncu_bench.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/r2_launches.csv
timeout 500 ncu --set full --clock-control none --import-source on -k regex:lstm_fwd_tcu -s 1 -c 1 -f -o gpurun_out/r2_prof_lstm_tcu python scripts/micro.py lstm_tcu > gpurun_out/r2_ncu_lstm_tcu.log 2>&1
echo "tcu prof exit $?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:lstm_bwd_tcu -s 1 -c 1 -f -o gpurun_out/r2_prof_lstm_tcu_bwd python scripts/micro.py lstm_tcu_bwd > gpurun_out/r2_ncu_lstm_tcu_bwd.log 2>&1
echo "tcu bwd prof exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ctc_loss_grad -s 2 -c 1 -f -o gpurun_out/r2_prof_ctc python scripts/micro.py ctc > gpurun_out/r2_ncu_ctc.log 2>&1
echo "ctc prof exit $?"
ls -la gpurun_out/r2_prof_*.ncu-rep
