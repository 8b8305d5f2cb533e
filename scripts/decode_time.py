"""Config-5 decoders (N=512, T=1000, C=22), inputs rotated over > L2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgr_b200 import ops
dev = torch.device("cuda:0")
N, T, C = 512, 1000, 22
bufs = [torch.softmax(torch.randn(N, T, C, device=dev) * 3, -1) for _ in range(6)]
i = [0]
def nxt():
    i[0] = (i[0] + 1) % len(bufs); return bufs[i[0]]
def timeit(fn, n=30):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, fn in (("bestpath_ref", lambda: ops.bestpath_ref(nxt(), 0.5)), ("greedy", lambda: ops.greedy(nxt()))):
    ms = timeit(fn)
    print("%s: %.1f us = %.0f GB/s" % (name, ms * 1e3, N * T * C * 4 / ms / 1e6))
