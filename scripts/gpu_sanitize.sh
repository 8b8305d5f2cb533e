mkdir -p gpurun_out
timeout 280 compute-sanitizer --tool memcheck python -m pytest "tests/test_gpu_lstm.py::test_every_recurrence_implementation" "tests/test_gpu_gemm.py::test_projection_variant_grouping" -q -m gpu -x > gpurun_out/sanitizer_memcheck.log 2>&1
echo "exit $?"; grep -v "Host Frame" gpurun_out/sanitizer_memcheck.log | head -60
