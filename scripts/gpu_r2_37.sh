mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -4
for bm in 1 0; do
GR_A32_BINMASK=$bm timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-ctc > gpurun_out/r2_bench_binmask$bm.json 2> gpurun_out/r2_bench_binmask$bm.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_binmask$bm.json").read().strip().splitlines()[-1])
print("BINMASK=$bm N=1 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "gemm", d["kernels"]["gr_gemm_a32_f32"])
PY
done
