"""First-layer projections through layers._project (bias in the padding column vs bias in the epilogue)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgr_b200 import ops, layers
dev = torch.device("cuda:0")
def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
T = 1000
for (BT, F, H) in ((256000, 39, 500), (256000, 20, 300), (32000, 39, 500), (32000, 20, 300)):
    x = torch.randn(BT, F, device=dev); W = torch.randn(F, 8 * H, device=dev) * 0.05; b = torch.randn(8 * H, device=dev)
    masks = ((torch.rand(8, BT // T, F, device=dev) > 0.5).float() * 2).contiguous()
    for bc in (True, False):
        layers.BIAS_COLUMN = bc
        ms = timed(lambda: layers._project(x, W, b, masks, BT // T, T, H))
        print("BT=%d F=%d H=%d bias column=%d: %.3f ms (whole _project incl. pad / split)" % (BT, F, H, bc, ms), flush=True)
    layers.BIAS_COLUMN = True
    g1 = layers._project(x, W, b, masks, BT // T, T, H)
    layers.BIAS_COLUMN = False
    g0 = layers._project(x, W, b, masks, BT // T, T, H)
    print("   max |diff| %.3e (scale %.2f)" % ((g1 - g0).abs().max().item(), g0.abs().max().item()))
