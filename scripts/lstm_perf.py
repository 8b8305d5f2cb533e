"""Recurrence-only timing (CUDA events) per layer shape, both arithmetic modes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgr_b200 import ops
dev = torch.device("cuda:0")
T = int(os.environ.get("T", "1000"))
shapes = [tuple(int(v) for v in x.split("x")) for x in os.environ.get("SHAPES", "256x500,256x300,128x500,32x500,32x300").split(",")]
for mode in ("bf16x3",):
    for (B, H) in shapes:
        gates = torch.randn(B * T, 8 * H, device=dev) * 0.5
        U = torch.randn(2, H, 4 * H, device=dev) / H ** 0.5
        for keep in (False, True):
            g2 = gates.clone()
            ops.lstm_recurrence_fwd(g2, U, B, T, H, keep_cell=keep)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g2 = gates.clone()
            e0.record()
            ops.lstm_recurrence_fwd(g2, U, B, T, H, keep_cell=keep)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print("precise=%s B=%d H=%d T=%d keep=%d: %.3f ms  %.2f us/step" % (mode, B, H, T, keep, ms, ms * 1e3 / T), flush=True)
        del gates, g2
