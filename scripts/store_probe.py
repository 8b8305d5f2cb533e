"""How fast can 4 GB be WRITTEN on this GPU (torch fill_), against the K=40 first-layer projection that is a pure store stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgr_b200 import ops, layers
dev = torch.device("cuda:0")
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
BT, T, F, H = 256000, 1000, 39, 500
out = torch.empty(BT, 8 * H, device=dev)
ms = timed(lambda: out.fill_(1.0))
print("fill_ of %.2f GB: %.3f ms = %.0f GB/s" % (out.numel() * 4 / 1e9, ms, out.numel() * 4 / ms / 1e6))
ms = timed(lambda: out.zero_())
print("zero_ (memset) : %.3f ms = %.0f GB/s" % (ms, out.numel() * 4 / ms / 1e6))
src = torch.randn(BT, 8 * H, device=dev)
ms = timed(lambda: out.copy_(src))
print("copy_: %.3f ms = %.0f GB/s read+write" % (ms, 2 * out.numel() * 4 / ms / 1e6))
del src
x = torch.randn(BT, F, device=dev); W = torch.randn(F, 8 * H, device=dev) * 0.05; b = torch.zeros(8 * H, device=dev)
masks = ((torch.rand(8, BT // T, F, device=dev) > 0.5).float() * 2).contiguous()
for env in ({}, {"GR_A32_EPI": "stg"}, {"GR_A32_EPI_BUFS": "6"}):
    for k, v in env.items(): os.environ[k] = v
    ms = timed(lambda: layers._project(x, W, b, masks, BT // T, T, H))
    print("first-layer projection K=39 N=4000 M=256000 %s: %.3f ms = %.0f GB/s written" % (env, ms, BT * 8 * H * 4 / ms / 1e6))
    for k in env: os.environ.pop(k)
