# First GPU call of round 2 (NEXT.md items 1 and 3), to be run with `gpurun --gpus 2`:
#   the not-yet-run CTC chunk instantiations, then the data-parallel pipeline with and without the candidate fix.
mkdir -p gpurun_out
GR_RUN_UNVERIFIED=1 timeout 200 python -m pytest tests/test_gpu_ctc.py -q -m gpu -k small_chunk --timeout 100 2>&1 | tail -4
for mode in 0 2 1; do
  echo "== N=2 GR_PIPELINE=$mode"
  GR_PIPELINE=$mode GR_BENCH_WATCHDOG_S=100 NCCL_DEBUG=WARN timeout 130 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
    --master-addr 127.0.0.1 --master-port $((29520 + mode)) bench.py --gpus 2 --steps 8 --warmup 3 --skip-cpu --skip-ctc \
    > gpurun_out/r2_dp2_p$mode.json 2> gpurun_out/r2_dp2_p$mode.err
  echo "exit $?"; tail -2 gpurun_out/r2_dp2_p$mode.err | cut -c1-200
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_dp2_p$mode.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "loss", d["loss_mean"])
except Exception as e:
    print("no result:", e)
PY
done
