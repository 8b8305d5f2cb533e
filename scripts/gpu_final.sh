mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 120 2>&1 | tail -4
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print("N=1 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "loss", d["loss_mean"], "launches", d["gpu_launches"])
for k, v in d["kernels"].items(): print("   ", k, v)
print(d["roofline"]); print(d["ctc"]); print(d["decode"]); print(d["cpu_baseline"]); print(d["clocks"])
PY
tail -3 gpurun_out/bench_final.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_final.err; tail -c 700 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches9.csv \
  python bench.py --steps 1 --warmup 1 --min-warmup 1 --skip-cpu --skip-ctc --skip-e2e > gpurun_out/ncu_bench9.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches9.csv
python -c "import __graft_entry__ as g; g.smoke()"
GR_PIPELINE=0 GR_TOWER_SPLIT=0 timeout 300 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-ctc > gpurun_out/bench_serial.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/bench_serial.json').read().strip().splitlines()[-1]); print('serial schedule (GR_PIPELINE=0 GR_TOWER_SPLIT=0):', d['value'], d['ms_per_step'], d['e2e']['value'])"
