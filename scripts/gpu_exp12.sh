for st in 1 0; do
GR_TOWER_STREAMS=$st timeout 300 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-ctc > gpurun_out/bench6_$st.json 2> gpurun_out/bench6_$st.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench6_$st.json").read().strip().splitlines()[-1])
print("streams=$st value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
for k, v in d["kernels"].items(): print("   ", k, v)
PY
done
GR_TOWER_STREAMS=0 timeout 100 python scripts/step_breakdown.py 2>&1 | tail -26
