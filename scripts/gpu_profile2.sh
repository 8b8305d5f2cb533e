mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches6.csv \
  python bench.py --steps 1 --warmup 1 --min-warmup 1 --skip-cpu --skip-ctc --skip-e2e > gpurun_out/ncu_bench6.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches6.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lstm_fwd_tc -s 1 -c 1 -f -o gpurun_out/prof_lstm_tc6 python scripts/micro.py lstm_tc > gpurun_out/ncu_lstm6.log 2>&1
echo "lstm prof exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_a32 -s 1 -c 1 -f -o gpurun_out/prof_a32_6 python scripts/micro.py a32 > gpurun_out/ncu_a32_6.log 2>&1
echo "a32 prof exit $?"
ls -la gpurun_out/*6.ncu-rep gpurun_out/prof_a32_6.ncu-rep
