mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err
echo "bench exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n1_b.json").read().strip().splitlines()[-1])
print("N=1 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "loss", d["loss_mean"], "launches", d["gpu_launches"])
for k, v in d["kernels"].items(): print("   ", k, v)
print(d["roofline"]); print({k: v for k, v in d["ctc"].items() if k != "sweep"}); print(d["decode"]); print(d["cpu_baseline"]); print(d["clocks"])
print("T1900", d["fusion_T1900"]); print("cfg1", d["config1_speech_fwd_loss"]["ms_per_step"], "cfg2", d["config2_skeletal_train"]["ms_per_step"], d["config2_skeletal_train"]["kernels"])
PY
tail -3 gpurun_out/r2_bench_n1_b.err
