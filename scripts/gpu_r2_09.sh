mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" GR_BENCH_WATCHDOG_S=200 timeout 260 python bench.py --steps 8 --warmup 3 --skip-cpu --skip-ctc > gpurun_out/r2_b_$name.json 2> gpurun_out/r2_b_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_b_$name.json").read().strip().splitlines()[-1])
    print("$name: value %.0f seq/s  %.2f ms/step  e2e %.0f  serial_step %.2f ms  loss %.6f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["serial_step_ms"], d["loss_mean"]))
except Exception as e:
    print("$name: no result:", e); print(open("gpurun_out/r2_b_$name.err").read()[-600:])
PY
}
run base A=1
run mainprio GR_MAIN_PRIO=-1
run towerlow GR_TOWER_PRIO=0 GR_MAIN_PRIO=-2
