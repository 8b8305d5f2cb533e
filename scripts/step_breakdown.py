"""Per-call breakdown of one fusion training step (CUDA events around every C-ABI call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgr_b200 as mgr
from mgr_b200 import _lib
import bench
dev = torch.device("cuda:0")
B, T = 256, 1000
model = mgr.FusionNet().to(dev); opt = mgr.fusion_optimizer(model)
xa, xs, lab, il, ll = [t.to(dev) for t in bench.synth_batch(0, B, T)]
def step(i):
    reg = model.sample_regularisers(B, T, seed=1, step=i, device=dev)
    loss, grads = model.loss_and_grads(xa, xs, lab, il, ll, reg)
    opt.step(grads)
for i in range(2): step(i)
names = list(_lib._SIGNATURES)
_lib.kernel_timing_begin(names)
order = []
orig = _lib.call
step(2)
t = _lib._timing
torch.cuda.synchronize()
rows = []
for n, evs in t.items():
    for k, (a, b) in enumerate(evs):
        rows.append((a, n, a.elapsed_time(b), _lib.kernel_timing_shapes[n][k]))
# order by start: use elapsed from first event
first = min(rows, key=lambda r: 0)[0]
base = None
rows2 = []
for a, n, ms, w in rows:
    rows2.append((rows[0][0].elapsed_time(a), n, ms, w))
rows2.sort()
tot = 0
for st, n, ms, w in rows2:
    if ms > 0.3:
        extra = ""
        if "gemm" in n and w: extra = "%.0f TFLOP/s alg" % (w / ms / 1e9)
        if "lstm" in n and w: extra = "%.0f GB/s alg" % (w / ms / 1e6)
        print("%8.2f ms  +%7.2f  %-30s %s" % (st, ms, n, extra))
    tot += ms
print("sum of calls %.1f ms" % tot)
