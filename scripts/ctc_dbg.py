import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import random_probs, random_labels
from mgr_b200 import ops
dev = torch.device("cuda:0")
B, T, C, Lmax = 8, 50, 22, 10
rng = np.random.default_rng(B * 1000 + T)
p, _ = random_probs(rng, B, T, C)
il = rng.integers((T - 2) // 2 + 1, T - 1, size=(B, 1)); il[0, 0] = T - 2
labels, ll = random_labels(rng, B, Lmax, C, T_avail=il[:, 0])
lab = torch.tensor(np.where(labels < 0, C - 1, labels).astype(np.int32), device=dev)
res = {}
for impl in ("v4", "v5"):
    os.environ["GR_CTC_IMPL"] = impl
    loss, grad, st = ops.ctc_loss_grad(torch.tensor(p, device=dev), lab, torch.tensor(ll[:, 0].astype(np.int32), device=dev),
                                       torch.tensor(il[:, 0].astype(np.int32), device=dev), False)
    res[impl] = (loss.cpu().numpy().copy(), grad.cpu().numpy().copy())
print("loss diff", np.abs(res["v4"][0] - res["v5"][0]).max())
g4, g5 = res["v4"][1], res["v5"][1]
for b in range(B):
    d = np.abs(g4[b] - g5[b]); sc = np.abs(g4[b]).max(axis=1, keepdims=True) + 1e-12
    rel = (d / sc).max(axis=1)
    bad = np.nonzero(rel > 1e-3)[0]
    print("b", b, "Tn", il[b, 0], "L", ll[b, 0], "labels", labels[b, :ll[b, 0]].astype(int).tolist(), "bad rows (t incl. 2 dropped):", bad.tolist()[:40])
    if len(bad):
        t = bad[0]
        print("   row", t, "v4", np.round(g4[b, t], 4).tolist())
        print("   row", t, "v5", np.round(g5[b, t], 4).tolist())
