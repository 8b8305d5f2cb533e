"""clock64 trace of the persistent projection GEMM (GR_A32_TRACE): per k-block stamps of the A producer,
the B TMA warp and the MMA warp, first 64 k-blocks of every CTA."""
import os, sys, ctypes
os.environ["GR_A32_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mgr_b200 import ops, layers, _lib
dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
if what in ("fwd", "ffwd", "l1"):
    BT, T, F, H = (65536, 1000, 1000, 500) if what == "fwd" else ((65536, 1000, 1600, 100) if what == "ffwd" else (65536, 1000, 40, 500))
    x = torch.randn(BT, F, device=dev); W = torch.randn(F, 8 * H, device=dev) * 0.05; b = torch.zeros(8 * H, device=dev)
    masks = ((torch.rand(8, BT // T + 1, F, device=dev) > 0.5).float() * 2).contiguous()
    run = lambda: layers._project(x, W, b, masks, BT // T, T, H)
else:
    BT, T, F, H = 65536, 1000, 1600, 100
    x = torch.randn(BT, F, device=dev); dP = torch.randn(BT, 8 * H, device=dev)
    masks = ((torch.rand(8, BT // T + 1, F, device=dev) > 0.5).float() * 2).contiguous()
    pt = ops.split_bf16(dP, transpose=True)
    dW = torch.empty(F, 8 * H, device=dev)
    run = lambda: ops.gemm_a32(x, pt[0], pt[1], F, H, BT, dW, 8 * H, nvar=8, mask=masks, rows_per_seq=T, transA=True)
for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print(what, "ms", e0.elapsed_time(e1))
lib = _lib.load()
n = 160 * 64 * 16
buf = (ctypes.c_longlong * n)()
assert lib.gr_debug_a32_trace(buf, ctypes.c_size_t(n)) == 0
tr = np.frombuffer(buf, dtype=np.int64).reshape(160, 64, 16)
tr = tr[:148]
names = ["prod: iteration start", "prod: stage empty", "prod: stored+arrived", "mma: fullB", "mma: fullA", "mma: issued+commit", "tmaB: stage empty(issue)", "-",
         "prod: half0 converted", "prod: half0 next loads issued", "prod: half1 converted", "prod: half1 next loads issued", "prod: fence done",
         "epi: tile begin (wait tmem_full)", "epi: tmem_full", "epi: tile done"]
base = tr[:, 20:60, 0:1]
rel = tr[:, 20:60, :16] - base
for i, nme in enumerate(names):
    print("  %-26s med %7d" % (nme, np.median(rel[:, :, i])))
per = np.median(np.diff(tr[:, 20:60, 2], axis=1))
print("  cycles per k-block (producer arrive to arrive): %d ; mma commit to commit: %d ; epilogue tile to tile: %d" % (per, np.median(np.diff(tr[:, 20:60, 5], axis=1)), np.median(np.diff(tr[:, 20:60, 15], axis=1))))
