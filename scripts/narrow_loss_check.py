"""Three FusionTrainer steps at the benchmark shape: losses with the narrow layer's forward on lstm_tcu (default at B=256)
and on lstm_small must agree (same Philox draws, same data)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgr_b200 as mgr
dev = torch.device("cuda:0")
B, T, C = 256, 1000, 22
def run():
    torch.manual_seed(0)
    model = mgr.FusionNet().to(dev)
    opt = mgr.fusion_optimizer(model)
    tr = mgr.FusionTrainer(model, opt, seed=99, global_batch=B)
    g = torch.Generator().manual_seed(5)
    xa = torch.randn(B, T, 39, generator=g).to(dev); xs = torch.randn(B, T, 20, generator=g).to(dev)
    rng = np.random.default_rng(6)
    labels = -np.ones((B, 40), dtype=np.float32); ll = np.zeros((B, 1), dtype=np.int64)
    for b in range(B):
        L = int(rng.integers(5, 41)); labels[b, :L] = rng.integers(0, C - 1, size=L); ll[b, 0] = L
    batch = (xa, xs, torch.tensor(labels), torch.tensor(np.full((B, 1), T - 2)), torch.tensor(ll))
    out = []
    for s in range(3):
        out.append(float(tr.step(batch, next_inputs=(xa, xs)).mean()))
    tr.close(); torch.cuda.synchronize()
    return out
a = run()
os.environ["GR_LSTM_NARROW_FWD"] = "small"
b = run()
print("narrow fwd on lstm_tcu :", a)
print("narrow fwd on lstm_small:", b)
print("max rel diff %.2e" % max(abs(x - y) / abs(y) for x, y in zip(a, b)))
