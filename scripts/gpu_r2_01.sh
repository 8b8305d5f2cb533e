# Round 2, call 1 (1 GPU): TS-mode MMA probe, the two CTC chunk instantiations that had never run, backward timing of the wide layers.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== ts_mma_probe"; timeout 120 scripts/micro/ts_mma_probe 2>&1 | tee gpurun_out/r2_ts_probe.log | head -60
echo "== CTC small chunks"
GR_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests/test_gpu_ctc.py -q -m gpu -k small_chunk --timeout 120 2>&1 | tail -6
GR_RUN_UNVERIFIED=1 timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ctc.py -q -m gpu -k small_chunk --timeout 250 2>&1 | tail -8 | tee gpurun_out/r2_sanitize_ctc_small.log
echo "== wide-layer backward (generic kernel) + forward"
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/r2_bwd_timing.log
import os, sys, torch
sys.path.insert(0, os.getcwd())
from mgr_b200 import ops
dev = torch.device("cuda:0")
for (B, T, H) in [(64, 800, 300), (16, 400, 500), (64, 800, 500)]:
    gates = torch.randn(B * T, 8 * H, device=dev) * 0.5
    U = torch.randn(2, H, 4 * H, device=dev) / H ** 0.5
    dy = torch.randn(B, T, 2 * H, device=dev) * 0.01
    for it in range(2):
        g2 = gates.clone()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        y, cell = ops.lstm_recurrence_fwd(g2, U, B, T, H, keep_cell=True)
        e[1].record()
        dP = ops.lstm_recurrence_bwd(g2, cell, dy, U, B, T, H)
        e[2].record(); torch.cuda.synchronize()
    print("B=%d T=%d H=%d: fwd(train) %.2f ms (%.2f us/step)  bwd %.2f ms (%.2f us/step)" % (B, T, H, e[0].elapsed_time(e[1]), e[0].elapsed_time(e[1]) * 1e3 / T, e[1].elapsed_time(e[2]), e[1].elapsed_time(e[2]) * 1e3 / T), flush=True)
PY
