mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --min-warmup 1 --skip-cpu --skip-ctc --skip-e2e > gpurun_out/ncu_bench.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctc_loss_grad -s 2 -c 1 -f -o gpurun_out/prof_ctc python scripts/micro.py ctc > gpurun_out/ncu_ctc.log 2>&1
echo "ctc prof exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_fwd_tc -s 1 -c 1 -f -o gpurun_out/prof_lstm_tc python scripts/micro.py lstm_tc > gpurun_out/ncu_lstm.log 2>&1
echo "lstm prof exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -f -o gpurun_out/prof_gemm python scripts/micro.py gemm > gpurun_out/ncu_gemm.log 2>&1
echo "gemm prof exit $?"
ls -la gpurun_out/*.ncu-rep
