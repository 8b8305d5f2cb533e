mkdir -p gpurun_out
for f in ctc decode gemm lstm models; do timeout 300 python -m pytest tests/test_gpu_$f.py -q -m gpu --timeout 100 2>&1 | tail -4; done
timeout 400 python bench.py --steps 3 --warmup 3 --skip-cpu > gpurun_out/bench4.json 2> gpurun_out/bench4.err
echo "bench exit $?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench4.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "loss", d["loss_mean"])
    for k, v in d["kernels"].items(): print(k, v)
    print("roofline", d["roofline"])
    print("ctc", d["ctc"]["ms"], d["ctc"]["roofline"]["frac"])
except Exception as e:
    print("no json", e)
PY
tail -5 gpurun_out/bench4.err
timeout 200 python scripts/step_breakdown.py 2>&1 | tail -40
