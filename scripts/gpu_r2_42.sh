mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches.csv \
  python bench.py --steps 1 --warmup 1 --min-warmup 1 --skip-cpu --skip-ctc --skip-e2e > gpurun_out/r2f_ncu_bench.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/r2f_launches.csv
bash scripts/gpu_r2_41.sh 2>&1 | grep -E "exit|^N=1|^\{.kernel|seq_len|config2|skeletal"
