mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_lstm.py -q -m gpu --timeout 60 -x 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --skip-cpu --skip-ctc > gpurun_out/bench_dp2.json 2> gpurun_out/bench_dp2.err
echo "dp2 exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_dp2.json").read().strip().splitlines()[-1])
print("N=2 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "loss", d["loss_mean"])
for k, v in d["kernels"].items(): print("   ", k, v)
PY
tail -3 gpurun_out/bench_dp2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dp_parity.py > gpurun_out/dp_parity.log 2>&1
echo "parity exit $?"; tail -5 gpurun_out/dp_parity.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench7.json 2> gpurun_out/bench7.err
echo "bench exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench7.json").read().strip().splitlines()[-1])
print("N=1 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "loss", d["loss_mean"])
for k, v in d["kernels"].items(): print("   ", k, v)
print(d["roofline"]); print(d["ctc"]); print(d["cpu_baseline"]); print(d["clocks"])
PY
tail -3 gpurun_out/bench7.err
