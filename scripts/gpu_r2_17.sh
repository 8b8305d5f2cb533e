mkdir -p gpurun_out
BS=256 T=400 timeout 300 ncu --set full --clock-control none --import-source on -k regex:lstm_small_fwd -s 1 -c 1 -f -o gpurun_out/r2_prof_small_fwd python scripts/small_perf.py time > gpurun_out/r2_ncu_small_fwd.log 2>&1
echo "prof exit $?"
BS=256 T=400 timeout 300 ncu --set full --clock-control none --import-source on -k regex:lstm_small_bwd -s 0 -c 1 -f -o gpurun_out/r2_prof_small_bwd python scripts/small_perf.py time > gpurun_out/r2_ncu_small_bwd.log 2>&1
echo "prof exit $?"
ls -la gpurun_out/r2_prof_small*.ncu-rep
