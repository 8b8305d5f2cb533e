mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_final.json 2>> gpurun_out/r2_bench_n1_final.err; echo "ref exit $?"; tail -c 600 gpurun_out/r2_bench_ref_final.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n1_final.json").read().strip().splitlines()[-1])
print("N=1 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
for k, v in d["kernels"].items(): print("   ", k, v)
print(d["roofline"]); print({k: d["ctc"][k] for k in d["ctc"] if k != "sweep"}); print(d["decode"]); print(d["cpu_baseline"]); print(d["clocks"])
print(d["fusion_T1900"]); print(d["config1_speech_fwd_loss"]); print(d["config2_skeletal_train"])
PY
