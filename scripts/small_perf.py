"""Narrow-layer recurrence (lstm_small.cu, H <= 104: the fusion BLSTM(100)) timing per batch size, forward
(inference / training) and BPTT, plus parity against the generic fp32 kernels on the same inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgr_b200 import ops
dev = torch.device("cuda:0")
T = int(os.environ.get("T", "1000"))
H = int(os.environ.get("H", "100"))
mode = sys.argv[1] if len(sys.argv) > 1 else "time"


def state(B, T, H, seed):
    gen = torch.Generator().manual_seed(seed)
    P = (torch.randn(B * T, 8 * H, generator=gen) * 0.7).to(dev)
    U = (torch.randn(2, H, 4 * H, generator=gen) / H ** 0.5).to(dev)
    dy = (torch.randn(B, T, 2 * H, generator=gen) * 0.1).to(dev)
    return P, U, dy


def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


if mode == "check":
    bad = 0
    for (B, T_, H_) in [(1, 5, 100), (3, 9, 100), (5, 17, 36), (32, 12, 100), (64, 7, 64), (130, 6, 100), (256, 8, 100), (37, 33, 104), (9, 40, 8)]:
        P, U, dy = state(B, T_, H_, B * 1000 + H_)
        os.environ["GR_LSTM_IMPL"] = "generic"
        g0 = P.clone(); y0, c0 = ops.lstm_recurrence_fwd(g0, U, B, T_, H_, keep_cell=True); y0 = y0.clone(); c0 = c0.clone()
        d0 = ops.lstm_recurrence_bwd(g0.clone(), c0, dy, U, B, T_, H_).clone()
        os.environ.pop("GR_LSTM_IMPL")
        g1 = P.clone(); y1, c1 = ops.lstm_recurrence_fwd(g1, U, B, T_, H_, keep_cell=True); y1 = y1.clone(); c1 = c1.clone()
        gi = P.clone(); yi, _ = ops.lstm_recurrence_fwd(gi, U, B, T_, H_, keep_cell=False); yi = yi.clone()
        d1 = ops.lstm_recurrence_bwd(g1.clone(), c1, dy, U, B, T_, H_).clone()
        d2 = ops.lstm_recurrence_bwd(g1.clone(), c1, dy, U, B, T_, H_).clone()
        torch.cuda.synchronize()
        ey = (y0 - y1).abs().max().item(); ec = (c0 - c1).abs().max().item(); eg = (g0 - g1).abs().max().item()
        ei = (y0 - yi).abs().max().item()
        ed = (d0 - d1).abs().max().item(); sd = d0.abs().max().item()
        ok = max(ey, ec, eg, ei) <= 2e-5 and ed <= 2e-4 * sd + 1e-7 and torch.equal(d1, d2)
        bad += not ok
        print(("ok   " if ok else "FAIL ") + "B=%d T=%d H=%d: y %.1e cell %.1e gates %.1e y(inference) %.1e dP %.1e (scale %.1e) det %s"
              % (B, T_, H_, ey, ec, eg, ei, ed, sd, torch.equal(d1, d2)), flush=True)
    print("small check:", "ALL OK" if bad == 0 else "%d FAILED" % bad)
else:
    for B in [int(v) for v in os.environ.get("BS", "8,32,64,128,256").split(",")]:
        P, U, dy = state(B, T, H, 1)
        g = P.clone()
        ms0 = timed(lambda: ops.lstm_recurrence_fwd(g, U, B, T, H, keep_cell=False))
        g = P.clone()
        ms1 = timed(lambda: ops.lstm_recurrence_fwd(g, U, B, T, H, keep_cell=True))
        g = P.clone(); y, cell = ops.lstm_recurrence_fwd(g, U, B, T, H, keep_cell=True); cell = cell.clone()
        g2 = g.clone()
        ms2 = timed(lambda: ops.lstm_recurrence_bwd(g2, cell, dy, U, B, T, H))
        print("small H=%d T=%d B=%d: fwd(inference) %.3f ms  fwd(train) %.3f ms  bwd %.3f ms  = %.2f / %.2f / %.2f us per step"
              % (H, T, B, ms0, ms1, ms2, ms0 * 1e3 / T, ms1 * 1e3 / T, ms2 * 1e3 / T), flush=True)
