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
T = 1000
for (BT, F, H) in ((256000, 40, 500), (256000, 20, 300), (32000, 40, 500)):
    x = torch.randn(BT, F, device=dev); W = torch.randn(F, 8 * H, device=dev) * 0.05; b = torch.zeros(8 * H, device=dev)
    masks = ((torch.rand(8, BT // T, F, device=dev) > 0.5).float() * 2).contiguous()
    gates = torch.empty((BT, 8 * H), dtype=torch.float32, device=dev)
    wt_hi, wt_lo = ops.split_bf16(W, transpose=True)
    fn = lambda: ops.gemm_a32(x, wt_hi, wt_lo, BT, H, F, gates, 8 * H, nvar=8, mask=masks, rows_per_seq=T, bias=b)
    for rep in range(3):
        for epi in ("tma", "wide"):
            os.environ["GR_A32_EPI"] = epi
            ms = timed(fn)
            print("BT=%d F=%d H=%d %s: %.3f ms = %.0f GB/s" % (BT, F, H, epi, ms, BT * 8 * H * 4 / ms / 1e6), flush=True)
