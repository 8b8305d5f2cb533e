mkdir -p gpurun_out
GR_BENCH_WATCHDOG_S=600 timeout 700 python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err
echo "exit $?"; tail -3 gpurun_out/r2_bench_full.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_full.json").read().strip().splitlines()[-1])
print("value %.0f  ms %.2f  e2e %.0f  launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
print("roofline", d["roofline"])
print("T1900", d["fusion_T1900"])
for k in ("config1_speech_fwd_loss", "config2_skeletal_train"):
    c = d[k]; print(k, {x: c[x] for x in c if x not in ("kernels",)}); print("   ", c["kernels"])
print("ctc", {k: v for k, v in d["ctc"].items() if k != "sweep"})
for r in d["ctc"]["sweep"]: print("   ", r)
print("decode", d["decode"])
print("cpu", d["cpu_baseline"])
PY
