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
    run = lambda: layers._project(x, W, b, masks, BT // T, T, H, mask_scale=float(os.environ.get('TRACE_MASK_SCALE', '0')))
else:
    BT, T, F, H = 65536, 1000, 1600, 100
    x = torch.randn(BT, F, device=dev); dP = torch.randn(BT, 8 * H, device=dev)
    masks = ((torch.rand(8, BT // T + 1, F, device=dev) > 0.5).float() * 2).contiguous()
    pt = ops.split_bf16(dP, transpose=True)
    dW = torch.empty(F, 8 * H, device=dev)
    run = lambda: ops.gemm_a32(x, pt[0], pt[1], F, H, BT, dW, 8 * H, nvar=8, mask=masks, rows_per_seq=T, transA=True, mask_scale=float(os.environ.get('TRACE_MASK_SCALE', '0')))
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

d = np.diff(tr[:, 8:60, 2], axis=1).astype(np.float64)
d = d[(tr[:, 9:60, 2] > 0) & (tr[:, 8:59, 2] > 0)]
print("  producer period over stages 8..59: mean %.0f  p10 %.0f  p50 %.0f  p90 %.0f  p99 %.0f  max %.0f" % (d.mean(), *np.percentile(d, [10, 50, 90, 99]), d.max()))
dd = np.diff(tr[:, 8:60, 2], axis=1).astype(np.float64)
for ph in range(4):
    sel = dd[:, ph::4]
    print("    stage %% 4 == %d: mean %.0f median %.0f" % ((8 + 1 + ph) % 4, sel.mean(), np.median(sel)))
w = (tr[:, 8:60, 1] - tr[:, 8:60, 0]).astype(np.float64)
print("  producer wait for the empty stage: mean %.0f p50 %.0f p90 %.0f" % (w.mean(), *np.percentile(w, [50, 90])))

# per-tile stamps (index = tile count of the CTA): MMA warp waiting for the drained accumulators, epilogue begin / tmem_full / done
tt = tr[:, 0:8, :]
ok = (tt[:, 1:7, 13] > 0) & (tt[:, 1:7, 15] > 0)
if ok.any():
    print("  per tile (tiles 1..6 of a CTA): epilogue tmem_full -> done  mean %.0f ; tile period (done to done) mean %.0f" % (
        (tt[:, 1:7, 15] - tt[:, 1:7, 14])[ok].mean(), np.diff(tt[:, 0:7, 15], axis=1)[ok].mean()))
    if (tt[:, 1:7, 11] > 0).any():
        okm = ok & (tt[:, 1:7, 11] > 0)
        print("  MMA warp: wait for drained accumulators mean %.0f ; first MMA of a tile after the last of the previous one: n/a" % ((tt[:, 1:7, 11] - tt[:, 1:7, 10])[okm].mean()))

# block-start anatomy of the dropout path (stamps at stage index = first stage of a k-block)
bs = tr[:, 8:60:4, :]
if (bs[:, :, 8] > 0).any():
    prev = tr[:, 7:59:4, 2]
    print("  block start: prev arrive -> ballots+mask loads issued %.0f -> barrier passed %.0f -> split done (x landed) %.0f -> x loads issued %.0f -> stage-0 arrive %.0f" % tuple(
        np.median(bs[:, :, sl] - prev) for sl in (7, 8, 9, 12, 2)))
