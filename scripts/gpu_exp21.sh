timeout 250 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_lstm.py tests/test_gpu_models.py -q -m gpu --timeout 60 -x 2>&1 | tail -4
for st in 1 0; do
GR_TOWER_STREAMS=$st timeout 200 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-ctc > gpurun_out/bench9_$st.json 2> gpurun_out/bench9_$st.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench9_$st.json").read().strip().splitlines()[-1])
print("streams=$st value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
done
timeout 100 python scripts/trace_a32.py fwd | grep "ms\|cycles per"
