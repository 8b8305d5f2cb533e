"""Config-2 layer (B=64, T=800, H=300): one recurrence launch vs two half-batch launches on two streams (38 + 38 CTAs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgr_b200 import ops
dev = torch.device("cuda:0")
B, T, H = 64, 800, 300
gates = torch.randn(B * T, 8 * H, device=dev) * 0.5
U = torch.randn(2, H, 4 * H, device=dev) / H ** 0.5
dy = torch.randn(B, T, 2 * H, device=dev) * 0.1
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def whole():
    g = gates.clone()
    y, c = ops.lstm_recurrence_fwd(g, U, B, T, H, keep_cell=True)
    ops.lstm_recurrence_bwd(g, c, dy, U, B, T, H)
def halves():
    g = gates.clone().reshape(B, T, 8 * H)
    cur = torch.cuda.current_stream()
    outs = []
    for i, s in enumerate((s1, s2)):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            gh = g[i * 32:(i + 1) * 32].reshape(32 * T, 8 * H)
            y, c = ops.lstm_recurrence_fwd(gh, U, 32, T, H, keep_cell=True)
            ops.lstm_recurrence_bwd(gh, c, dy[i * 32:(i + 1) * 32].contiguous(), U, 32, T, H)
            outs.append((y, c))
    for s in (s1, s2): cur.wait_stream(s)
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for rep in range(2):
    print("whole batch fwd(train)+bwd: %.2f ms   two halves on two streams: %.2f ms" % (timed(whole), timed(halves)), flush=True)
