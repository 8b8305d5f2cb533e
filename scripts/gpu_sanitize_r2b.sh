# compute-sanitizer over the kernels changed in the last session of round 2: projection GEMM (dropout producers, keep-bit ring,
# per-variant accumulator hand-off, producer-drained store-stream epilogue), auxiliary output of the recurrence, dense backward,
# 64x64 transposing split
mkdir -p gpurun_out
T1="tests/test_gpu_gemm.py::test_fused_prologue_projection tests/test_gpu_gemm.py::test_fused_prologue_weight_gradient tests/test_gpu_gemm.py::test_projection_epilogue_variants"
T2="tests/test_gpu_lstm.py::test_recurrence_auxiliary_output tests/test_gpu_lstm.py::test_tower_residual_through_recurrence tests/test_gpu_models.py::test_dense_softmax_kernels"
timeout 900 compute-sanitizer --tool memcheck python -m pytest $T1 $T2 -q -m gpu -x --timeout 850 > gpurun_out/r2b_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"; grep -v "Host Frame" gpurun_out/r2b_sanitizer_memcheck.log | tail -6
timeout 900 compute-sanitizer --tool racecheck python -m pytest $T1 -q -m gpu -x -k "1000-200-1600 or 768-128-40 or 3000-1000 or 4224-66 or epilogue" --timeout 850 > gpurun_out/r2b_sanitizer_racecheck.log 2>&1
echo "racecheck exit $?"; grep -v "Host Frame" gpurun_out/r2b_sanitizer_racecheck.log | tail -6
timeout 600 compute-sanitizer --tool synccheck python -m pytest $T1 tests/test_gpu_lstm.py::test_recurrence_auxiliary_output -q -m gpu -x -k "1000-200-1600 or 768-128-40 or 3000-1000 or auxiliary" --timeout 550 > gpurun_out/r2b_sanitizer_synccheck.log 2>&1
echo "synccheck exit $?"; grep -v "Host Frame" gpurun_out/r2b_sanitizer_synccheck.log | tail -6
