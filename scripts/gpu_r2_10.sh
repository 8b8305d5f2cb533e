mkdir -p gpurun_out
GR_TCU_NB=64 timeout 300 python scripts/tcu_check.py time 2>&1 | grep -E "tcu B=256|tcu B=128" 
GR_TCU_NB=64 GR_BENCH_WATCHDOG_S=200 timeout 260 python bench.py --steps 8 --warmup 3 --skip-cpu --skip-ctc > gpurun_out/r2_b_nb64.json 2> gpurun_out/r2_b_nb64.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_b_nb64.json").read().strip().splitlines()[-1])
print("nb64: value %.0f seq/s  %.2f ms/step  e2e %.0f  serial_step %.2f ms" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["serial_step_ms"]))
PY
