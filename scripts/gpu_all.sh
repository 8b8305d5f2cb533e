mkdir -p gpurun_out
for f in ctc decode gemm lstm models; do timeout 300 python -m pytest tests/test_gpu_$f.py tests/test_golden.py -q -m gpu --timeout 100 2>&1 | tail -3; done
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench8.json 2> gpurun_out/bench8.err
echo "bench exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench8.json").read().strip().splitlines()[-1])
print("N=1 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "loss", d["loss_mean"], "launches", d["gpu_launches"])
for k, v in d["kernels"].items(): print("   ", k, v)
print(d["roofline"]); print(d["ctc"]); print(d["decode"]); print(d["cpu_baseline"]); print(d["clocks"])
PY
tail -3 gpurun_out/bench8.err
timeout 250 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_lstm.py::test_every_recurrence_implementation" "tests/test_gpu_gemm.py::test_projection_variant_grouping" -q -m gpu -x 2>&1 | tail -6
echo "sanitizer exit $?"
