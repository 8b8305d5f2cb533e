# session 3 baseline: full GPU suite + smoke + default bench at HEAD
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -8 | tee gpurun_out/r2_gpu_tests_s3.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_bench_n1_s3.json 2> gpurun_out/r2_bench_n1_s3.err
echo "bench exit $?"; tail -c 400 gpurun_out/r2_bench_n1_s3.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n1_s3.json").read().strip().splitlines()[-1])
print("N=1 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
for k, v in d["kernels"].items(): print("   ", k, v)
print({k: d["ctc"][k] for k in d["ctc"] if k != "sweep"})
PY
