timeout 250 python -m pytest tests/test_gpu_models.py -q -m gpu --timeout 60 -x 2>&1 | tail -4
timeout 200 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-ctc > gpurun_out/bench10.json 2> gpurun_out/bench10.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench10.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "loss", d["loss_mean"])
PY
tail -3 gpurun_out/bench10.err
timeout 100 python scripts/step_breakdown.py 2>&1 | tail -24
