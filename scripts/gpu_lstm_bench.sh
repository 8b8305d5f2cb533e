mkdir -p gpurun_out
for f in lstm models; do timeout 300 python -m pytest tests/test_gpu_$f.py -q -m gpu --timeout 120 2>&1 | tail -12; done
timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu > gpurun_out/bench2.json 2> gpurun_out/bench2.err
echo "bench exit $?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench2.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
    for k, v in d["kernels"].items(): print(k, v)
    print("roofline", d["roofline"]); print("ctc", d["ctc"])
except Exception as e:
    print("no json", e)
PY
tail -5 gpurun_out/bench2.err
