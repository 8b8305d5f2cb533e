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
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches8.csv \
  python bench.py --steps 1 --warmup 1 --min-warmup 1 --skip-cpu --skip-ctc --skip-e2e > gpurun_out/ncu_bench8.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches8.csv
python -c "import __graft_entry__ as g; g.smoke()"
timeout 280 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_ctc.py -q -m gpu -x -k "chunk_boundaries and v4 and (33 or 65 or 129)" > gpurun_out/sanitizer_racecheck_ctc.log 2>&1
echo "racecheck exit $?"; grep -v "Host Frame" gpurun_out/sanitizer_racecheck_ctc.log | tail -6
timeout 280 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ctc.py -q -m gpu -x -k "chunk_boundaries and v4 or edge_cases or 300-22-150" > gpurun_out/sanitizer_memcheck_ctc.log 2>&1
echo "memcheck exit $?"; grep -v "Host Frame" gpurun_out/sanitizer_memcheck_ctc.log | tail -6
