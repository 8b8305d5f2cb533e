"""lstm_tcu.cu (U in tensor memory) against the generic fp32 recurrence and the round-1 tcgen05 kernel:
outputs, saved gates / cell state, run-to-run determinism; then timing per layer shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgr_b200 import ops
dev = torch.device("cuda:0")


def run(impl, gates, U, B, T, H, keep):
    os.environ["GR_LSTM_IMPL"] = impl
    g = gates.clone()
    y, cell = ops.lstm_recurrence_fwd(g, U, B, T, H, keep_cell=keep)
    torch.cuda.synchronize()
    return y, cell, g


mode = sys.argv[1] if len(sys.argv) > 1 else "check"
if mode == "check":
    shapes = [(5, 9, 36), (16, 7, 64), (40, 12, 300), (130, 6, 100), (64, 10, 500), (32, 10, 300), (200, 7, 64),
              (256, 20, 500), (256, 12, 300), (512, 5, 300), (37, 33, 500)]
    bad = 0
    for (B, T, H) in shapes:
        gen = torch.Generator().manual_seed(B * 1000 + H)
        gates = (torch.randn(B * T, 8 * H, generator=gen) * 0.7).to(dev)
        U = (torch.randn(2, H, 4 * H, generator=gen) / H ** 0.5).to(dev)
        for keep in (False, True):
            y0, c0, g0 = run("generic", gates, U, B, T, H, keep)
            y1, c1, g1 = run("tcu", gates, U, B, T, H, keep)
            y2, _, _ = run("tcu", gates, U, B, T, H, keep)
            ey = (y0 - y1).abs().max().item()
            msg = "B=%d T=%d H=%d keep=%d: |y - generic| %.2e  deterministic %s" % (B, T, H, keep, ey, torch.equal(y1, y2))
            ok = ey <= 2e-4 and torch.equal(y1, y2)
            if keep:
                ec, eg = (c0 - c1).abs().max().item(), (g0 - g1).abs().max().item()
                msg += "  |c| %.2e |gates| %.2e" % (ec, eg)
                ok = ok and ec <= 5e-4 and eg <= 2e-4
            print(("ok   " if ok else "FAIL ") + msg, flush=True)
            bad += (not ok)
            if not ok and ey > 1e-2:
                d = (y0 - y1).abs()
                idx = torch.nonzero(d > 1e-2)
                print("   first bad (b,t,col):", idx[:6].tolist(), " n_bad", idx.shape[0], flush=True)
    print("tcu check:", "ALL OK" if bad == 0 else "%d FAILED" % bad)
else:
    T = int(os.environ.get("T", "1000"))
    for (B, H) in [(256, 500), (256, 300), (128, 500), (64, 500), (32, 500), (64, 300), (32, 300)]:
        gates = torch.randn(B * T, 8 * H, device=dev) * 0.5
        U = torch.randn(2, H, 4 * H, device=dev) / H ** 0.5
        for impl in ("tc", "tcu"):
            for keep in (False, True):
                os.environ["GR_LSTM_IMPL"] = impl
                g2 = gates.clone()
                ops.lstm_recurrence_fwd(g2, U, B, T, H, keep_cell=keep)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g2 = gates.clone()
                e0.record()
                ops.lstm_recurrence_fwd(g2, U, B, T, H, keep_cell=keep)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                print("%-3s B=%d H=%d T=%d keep=%d: %.3f ms  %.2f us/step" % (impl, B, H, T, keep, ms, ms * 1e3 / T), flush=True)
        del gates
