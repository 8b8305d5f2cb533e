"""Fusion BLSTM(100) at B=256: tensor-memory kernels (forward / BPTT, tile widths) against the register-resident kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgr_b200 import ops
dev = torch.device("cuda:0")
B, T, H = int(os.environ.get("B", "256")), 1000, 100
gates = torch.randn(B * T, 8 * H, device=dev) * 0.5
U = torch.randn(2, H, 4 * H, device=dev) / H ** 0.5
dy = torch.randn(B, T, 2 * H, device=dev) * 0.1
def run(impl, fnb=None, bnb=None):
    for k in ("GR_LSTM_IMPL", "GR_TCU_NB", "GR_TCU_BWD_NB"): os.environ.pop(k, None)
    if impl: os.environ["GR_LSTM_IMPL"] = impl
    if fnb: os.environ["GR_TCU_NB"] = str(fnb)
    if bnb: os.environ["GR_TCU_BWD_NB"] = str(bnb)
    g = gates.clone(); y, c = ops.lstm_recurrence_fwd(g, U, B, T, H, keep_cell=True); ops.lstm_recurrence_bwd(g, c, dy, U, B, T, H)
    torch.cuda.synchronize()
    g = gates.clone()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); y, c = ops.lstm_recurrence_fwd(g, U, B, T, H, keep_cell=True); e[1].record()
    dP = ops.lstm_recurrence_bwd(g, c, dy, U, B, T, H); e[2].record(); torch.cuda.synchronize()
    nf, nbk = ops.lstm_recurrence_grid(B, H), 0
    print("impl=%-6s fwd NB=%-4s bwd NB=%-4s: fwd(train) %.2f ms  bwd %.2f ms" % (impl or "small", fnb, bnb, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])), flush=True)
    return y.clone(), dP.clone()
y0, d0 = run(None)
for fnb, bnb in ((128, 128), (64, 64), (32, 32), (16, 16)):
    y1, d1 = run("tcu", fnb, bnb)
    print("    |y - small| %.2e  |dP - small| %.2e (scale %.2e)" % ((y1 - y0).abs().max().item(), (d1 - d0).abs().max().item(), d0.abs().max().item()))
