# final round-2 evidence: ncu launch list of the bench command + `--set full` captures of the current kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches.csv \
  python bench.py --steps 1 --warmup 1 --min-warmup 1 --skip-cpu --skip-ctc --skip-e2e > gpurun_out/r2f_ncu_bench.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/r2f_launches.csv
cap() { # name regex workload skip
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -f -o gpurun_out/r2f_prof_$1 python scripts/micro.py $3 > gpurun_out/r2f_ncu_$1.log 2>&1; echo "$1 exit $?"; }
cap ctc ctc_loss_grad ctc 2
cap a32_k1000 gemm_a32 a32full 1
cap lstm_tcu lstm_fwd_tcu lstm_tcu 1
cap small_fwd lstm_small_fwd small 1
cap small_bwd lstm_small_bwd small 1
ls -la gpurun_out/r2f_prof_*.ncu-rep
