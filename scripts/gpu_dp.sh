# data-parallel runs on N GPUs of one box: `gpurun --gpus N -- 'N=2 REPS=3 bash scripts/gpu_dp.sh'`
mkdir -p gpurun_out
N=${N:-2}; REPS=${REPS:-2}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29610 scripts/dp_parity.py 2>&1 | grep -E "rank|Error|error" | head -12
for i in $(seq 1 $REPS); do
  GR_BENCH_WATCHDOG_S=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29620 + i)) \
    bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --skip-cpu --skip-ctc > gpurun_out/r2_dp${N}_$i.json 2> gpurun_out/r2_dp${N}_$i.err
  echo "N=$N run $i exit $?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_dp${N}_$i.json").read().strip().splitlines()[-1])
    print("  value %.0f seq/s  %.2f ms/step  e2e %.0f  serial %.2f ms  loss %.6f  schedule: %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["serial_step_ms"], d["loss_mean"], d["config"]["schedule"][:60]))
except Exception as e:
    print("  no result:", e)
PY
  tail -2 gpurun_out/r2_dp${N}_$i.err | cut -c1-200
done
