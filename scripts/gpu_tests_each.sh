mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 300 python scripts/debug_lstm.py > gpurun_out/debug_lstm.log 2>&1
for f in ctc decode gemm lstm models; do
  timeout 420 python -m pytest tests/test_gpu_$f.py -q -m gpu --timeout 200 > gpurun_out/t_$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  tail -3 gpurun_out/t_$f.log
done
cat gpurun_out/summary.txt
cat gpurun_out/debug_lstm.log | tail -20
