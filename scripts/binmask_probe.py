"""Fusion-layer projection (row-major A, K=1600, 8 x 100 columns) and weight gradient (transposed A, K = B*T) with the
masks declared as dropout masks (shared split, keep bits) against the generic multiply path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgr_b200 import ops
dev = torch.device("cuda:0")
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
T, F, H = 1000, 1600, 100
for BT in (256000, 32000):
    x = torch.randn(BT, F, device=dev); W = torch.randn(F, 8 * H, device=dev) * 0.05; b = torch.zeros(8 * H, device=dev)
    dP = torch.randn(BT, 8 * H, device=dev)
    masks = ((torch.rand(8, BT // T, F, device=dev) > 0.5).float() * 2).contiguous()
    gates = torch.empty((BT, 8 * H), dtype=torch.float32, device=dev)
    dW = torch.empty(F, 8 * H, device=dev)
    wt = ops.split_bf16(W, transpose=True)
    pt = ops.split_bf16(dP, transpose=True)
    for rep in range(2):
        for name, env, ms_ in (("generic nvg=1", {}, 0.0), ("generic nvg=4", {"GR_A32_NVG": "4"}, 0.0), ("dropout nvg=4", {}, 2.0)):
            os.environ.update(env)
            ms = timed(lambda: ops.gemm_a32(x, wt[0], wt[1], BT, H, F, gates, 8 * H, nvar=8, mask=masks, rows_per_seq=T, bias=b, mask_scale=ms_))
            print("BT=%d projection %-14s %.3f ms = %.0f TFLOP/s" % (BT, name, ms, 2.0 * BT * F * 8 * H / ms / 1e9), flush=True)
            for k in env: os.environ.pop(k)
        for name, ms_ in (("generic nvg=4", 0.0), ("dropout nvg=4", 2.0)):
            ms = timed(lambda: ops.gemm_a32(x, pt[0], pt[1], F, H, BT, dW, 8 * H, nvar=8, mask=masks, rows_per_seq=T, transA=True, mask_scale=ms_))
            print("BT=%d dW         %-14s %.3f ms = %.0f TFLOP/s" % (BT, name, ms, 2.0 * BT * F * 8 * H / ms / 1e9), flush=True)
