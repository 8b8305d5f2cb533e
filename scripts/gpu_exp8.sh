for f in gemm lstm models; do timeout 200 python -m pytest tests/test_gpu_$f.py -q -m gpu --timeout 60 -x 2>&1 | tail -15; done
timeout 300 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-ctc > gpurun_out/bench5.json 2> gpurun_out/bench5.err
echo "bench exit $?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench5.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "loss", d["loss_mean"])
    for k, v in d["kernels"].items(): print(k, v)
except Exception as e:
    print("no json", e)
PY
tail -5 gpurun_out/bench5.err
timeout 100 python scripts/step_breakdown.py 2>&1 | tail -30
