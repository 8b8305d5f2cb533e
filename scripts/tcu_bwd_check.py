"""lstm_tcu.cu backward (tensor-core BPTT) against the generic fp32 kernel on the same saved forward state; timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgr_b200 import ops
dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "check"


def fwd_state(B, T, H, seed):
    gen = torch.Generator().manual_seed(seed)
    P = (torch.randn(B * T, 8 * H, generator=gen) * 0.7).to(dev)
    U = (torch.randn(2, H, 4 * H, generator=gen) / H ** 0.5).to(dev)
    dy = (torch.randn(B, T, 2 * H, generator=gen) * 0.1).to(dev)
    os.environ["GR_LSTM_IMPL"] = "generic"
    g = P.clone()
    y, cell = ops.lstm_recurrence_fwd(g, U, B, T, H, keep_cell=True)
    torch.cuda.synchronize()
    return g, cell, dy, U


if mode == "check":
    bad = 0
    for (B, T, H) in [(5, 9, 36), (16, 7, 64), (40, 12, 300), (64, 10, 300), (130, 6, 100), (64, 10, 500), (200, 7, 64),
                      (256, 8, 500), (37, 33, 500)]:
        g, cell, dy, U = fwd_state(B, T, H, B * 1000 + H)
        os.environ["GR_LSTM_IMPL"] = "generic"
        d0 = ops.lstm_recurrence_bwd(g.clone(), cell, dy, U, B, T, H).clone()
        os.environ["GR_LSTM_IMPL"] = "tcu"
        d1 = ops.lstm_recurrence_bwd(g.clone(), cell, dy, U, B, T, H).clone()
        d2 = ops.lstm_recurrence_bwd(g.clone(), cell, dy, U, B, T, H).clone()
        torch.cuda.synchronize()
        scale = d0.abs().max().item()
        err = (d0 - d1).abs().max().item()
        ok = err <= 2e-4 * max(scale, 1e-6) + 1e-7 and torch.equal(d1, d2)
        print(("ok   " if ok else "FAIL ") + "B=%d T=%d H=%d: |dP - generic| %.2e (scale %.2e)  deterministic %s" % (B, T, H, err, scale, torch.equal(d1, d2)), flush=True)
        if not ok:
            bad += 1
            d = (d0 - d1).abs().reshape(B, T, 8 * H)
            idx = torch.nonzero(d > 0.01 * scale)
            print("   first bad (b,t,col):", idx[:6].tolist(), " n_bad", idx.shape[0], flush=True)
    print("tcu bwd check:", "ALL OK" if bad == 0 else "%d FAILED" % bad)
else:
    T = int(os.environ.get("T", "800"))
    for (B, H) in [(64, 300), (64, 500), (16, 500), (256, 300), (256, 500)]:
        g, cell, dy, U = fwd_state(B, T, H, 1)
        for impl in ("tcu", "generic"):
            if impl == "generic" and B > 64:
                continue
            os.environ["GR_LSTM_IMPL"] = impl
            ops.lstm_recurrence_bwd(g.clone(), cell, dy, U, B, T, H)
            torch.cuda.synchronize()
            g2 = g.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.lstm_recurrence_bwd(g2, cell, dy, U, B, T, H)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print("%-7s bwd B=%d H=%d T=%d: %.3f ms  %.2f us/step" % (impl, B, H, T, ms, ms * 1e3 / T), flush=True)
