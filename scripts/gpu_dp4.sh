mkdir -p gpurun_out
GR_BENCH_WATCHDOG_S=120 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 5 --warmup 3 --skip-cpu --skip-ctc > gpurun_out/bench_dp4.json 2> gpurun_out/bench_dp4.err
echo "dp4 exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_dp4.json").read().strip().splitlines()[-1])
print("N=4 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "loss", d["loss_mean"])
PY
tail -3 gpurun_out/bench_dp4.err
