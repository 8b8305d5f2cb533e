mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 5 --warmup 3 --skip-cpu --skip-ctc > gpurun_out/bench_dp4.json 2> gpurun_out/bench_dp4.err
echo "dp4 exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_dp4.json").read().strip().splitlines()[-1])
print("N=4 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "loss", d["loss_mean"])
for k, v in d["kernels"].items(): print("   ", k, v)
PY
tail -3 gpurun_out/bench_dp4.err
