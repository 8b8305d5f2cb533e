N=${N:-2}
mkdir -p gpurun_out
GR_BENCH_WATCHDOG_S=240 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 --skip-cpu --skip-ctc > gpurun_out/r2_bench_dp${N}_d.json 2> gpurun_out/r2_bench_dp${N}_d.err
echo "dp$N exit $?"; python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_dp${N}_d.json").read().strip().splitlines()[-1])
print("N=$N value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "loss", d["loss_mean"], "serial", d["roofline"]["serial_step_ms"])
for k, v in d["kernels"].items(): print("   ", k, v)
PY
tail -2 gpurun_out/r2_bench_dp${N}_d.err
