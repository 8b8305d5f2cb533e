"""Fusion-head kernels at the benchmark shape (B*T = 256000): dense backward, dense+softmax forward, dP^T split, colsum."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgr_b200 import ops
dev = torch.device("cuda:0")
def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
BT, Fin, C, H = 256000, 200, 22, 100
x = torch.randn(BT, Fin, device=dev); Wd = torch.randn(Fin, C, device=dev) * 0.1; bd = torch.zeros(C, device=dev)
m = ((torch.rand(BT, Fin, device=dev) > 0.5).float() * 2)
g = torch.randn(BT, C, device=dev)
dP = torch.randn(BT, 8 * H, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)   # 256 MB > L2
def fl(fn):
    def f():
        flush.zero_(); fn()
    return f
z = timed(lambda: flush.zero_())
for name, fn, bytes_ in (("dense_bwd", lambda: ops.dense_bwd(x, Wd, g, m, want_dx=True), BT * (Fin * 12 + C * 4)),
                         ("dense_softmax_fwd", lambda: ops.dense_softmax_fwd(x, Wd, bd, m, want_probs=False), BT * (Fin * 8 + C * 4)),
                         ("split_bf16 dP^T", lambda: ops.split_bf16(dP, transpose=True), BT * 8 * H * 8),
                         ("colsum dP", lambda: ops.colsum(dP), BT * 8 * H * 4)):
    ms = timed(fl(fn)) - z
    print("%-18s %.3f ms = %.0f GB/s" % (name, ms, bytes_ / ms / 1e6), flush=True)
