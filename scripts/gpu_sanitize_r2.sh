# compute-sanitizer passes over the two round-2 recurrence kernels (parity tests that force GR_LSTM_IMPL=tcu)
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool racecheck python -m pytest "tests/test_gpu_lstm.py::test_every_recurrence_implementation" -q -m gpu -x -k "tcu" --timeout 450 > gpurun_out/r2_sanitizer_racecheck_tcu.log 2>&1
echo "racecheck exit $?"; grep -v "Host Frame" gpurun_out/r2_sanitizer_racecheck_tcu.log | tail -12
timeout 500 compute-sanitizer --tool memcheck python -m pytest "tests/test_gpu_lstm.py::test_every_recurrence_implementation" "tests/test_gpu_lstm.py::test_tensor_core_bptt_matches_generic_and_is_deterministic" -q -m gpu -x -k "tcu or bptt" --timeout 450 > gpurun_out/r2_sanitizer_memcheck_tcu.log 2>&1
echo "memcheck exit $?"; grep -v "Host Frame" gpurun_out/r2_sanitizer_memcheck_tcu.log | tail -12
timeout 300 compute-sanitizer --tool synccheck python -m pytest "tests/test_gpu_lstm.py::test_every_recurrence_implementation" -q -m gpu -x -k "tcu-5-9" --timeout 250 > gpurun_out/r2_sanitizer_synccheck_tcu.log 2>&1
echo "synccheck exit $?"; grep -v "Host Frame" gpurun_out/r2_sanitizer_synccheck_tcu.log | tail -8
