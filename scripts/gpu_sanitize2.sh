mkdir -p gpurun_out
timeout 200 python -m pytest "tests/test_gpu_lstm.py::test_full_size_recurrence_properties" -q -m gpu -x 2>&1 | tail -5
timeout 280 compute-sanitizer --tool racecheck python -m pytest "tests/test_gpu_lstm.py::test_every_recurrence_implementation" -q -m gpu -x -k "tc" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck exit $?"; grep -v "Host Frame" gpurun_out/sanitizer_racecheck.log | tail -25
