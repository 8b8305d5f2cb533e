mkdir -p gpurun_out
for w in ffwd dw; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_a32 -s 2 -c 1 -f -o gpurun_out/r2_prof_a32_${w}_dropout python scripts/a32_one.py $w 2 > gpurun_out/r2_ncu_a32_${w}.log 2>&1; echo "exit $?"
done
