"""One fusion-layer projection / weight-gradient launch for ncu: python scripts/a32_one.py {ffwd|dw} [mask_scale]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgr_b200 import ops
dev = torch.device("cuda:0")
what = sys.argv[1]; ms_ = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
BT, T, F, H = 256000, 1000, 1600, 100
x = torch.randn(BT, F, device=dev); W = torch.randn(F, 8 * H, device=dev) * 0.05; b = torch.zeros(8 * H, device=dev)
masks = ((torch.rand(8, BT // T, F, device=dev) > 0.5).float() * 2).contiguous()
if what == "ffwd":
    gates = torch.empty((BT, 8 * H), dtype=torch.float32, device=dev)
    wt = ops.split_bf16(W, transpose=True)
    fn = lambda: ops.gemm_a32(x, wt[0], wt[1], BT, H, F, gates, 8 * H, nvar=8, mask=masks, rows_per_seq=T, bias=b, mask_scale=ms_)
else:
    dP = torch.randn(BT, 8 * H, device=dev); dW = torch.empty(F, 8 * H, device=dev)
    pt = ops.split_bf16(dP, transpose=True)
    fn = lambda: ops.gemm_a32(x, pt[0], pt[1], F, H, BT, dW, 8 * H, nvar=8, mask=masks, rows_per_seq=T, transA=True, mask_scale=ms_)
for _ in range(3): fn()
torch.cuda.synchronize()
