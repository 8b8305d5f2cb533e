import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import mgr_b200 as mgr
from mgr_b200 import ops
from oracle import lstm_ref
dev = torch.device("cuda:0")
t64 = lambda ws: [torch.tensor(w, dtype=torch.float64) for w in ws]

print("=== unimodal stage-wise")
rng = np.random.default_rng(1001)
B, T, F, H, C = 4, 40, 39, 24, 21
net = mgr.UnimodalNet(F, H, C, 0.5, (0.4, 0.5, 0.5), seed=5).to(dev)
x = rng.standard_normal((B, T, F)).astype(np.float32)
xt = torch.tensor(x, device=dev)
with torch.no_grad():
    y1 = net.blstm_1(xt); y2 = net.blstm_2(y1); res = ops.add(y1, y2)
    p, a = net.dense(res)
x64 = torch.tensor(x, dtype=torch.float64)
r1 = lstm_ref.bidirectional_lstm(x64, t64(net.blstm_1.get_weights()))
r2 = lstm_ref.bidirectional_lstm(r1, t64(net.blstm_2.get_weights()))
r2b = lstm_ref.bidirectional_lstm(y1.cpu().double(), t64(net.blstm_2.get_weights()))
rres = r1 + r2
rp, ra = lstm_ref.dense_softmax(rres, *t64(net.dense.get_weights()))
rp2, ra2 = lstm_ref.dense_softmax(res.cpu().double(), *t64(net.dense.get_weights()))
f = lambda a_, b_: float((a_.cpu().double() - b_).abs().max())
print("y1", f(y1, r1), "y2", f(y2, r2), "y2|gpu y1", f(y2, r2b), "res", f(res, rres), "logits", f(a, ra), "probs", f(p, rp),
      "dense only logits", f(a, ra2), "probs", f(p, rp2))
e = (y2.cpu().double() - r2b).abs()
print("y2 err by dir fwd/bwd", float(e[..., :H].max()), float(e[..., H:].max()), "argmax", np.unravel_index(int(e.argmax()), e.shape))

print("=== H=500 grads")
for masked in (False, True):
    rng = np.random.default_rng(6 * 100 + 8 + 500)
    B, T, F, H = 6, 8, 40, 500
    x = rng.standard_normal((B, T, F)).astype(np.float32)
    W = rng.uniform(-0.3, 0.3, size=(F, 8 * H)).astype(np.float32)
    U = (rng.standard_normal((2, H, 4 * H)) / np.sqrt(H)).astype(np.float32)
    b = (rng.standard_normal(8 * H) * 0.2).astype(np.float32)
    dy = rng.standard_normal((B, T, 2 * H)).astype(np.float32)
    masks = ((rng.random((8, B, F)) > 0.5) / 0.5).astype(np.float32) if masked else None
    xt = torch.tensor(x, device=dev, requires_grad=True); Wt = torch.tensor(W, device=dev, requires_grad=True)
    Ut = torch.tensor(U, device=dev, requires_grad=True); bt = torch.tensor(b, device=dev, requires_grad=True)
    mt = None if masks is None else torch.tensor(masks, device=dev)
    y = mgr.blstm(xt, Wt, Ut, bt, mt); y.backward(torch.tensor(dy, device=dev)); torch.cuda.synchronize()
    x6 = torch.tensor(x, dtype=torch.float64, requires_grad=True); W6 = torch.tensor(W, dtype=torch.float64, requires_grad=True)
    U6 = torch.tensor(U, dtype=torch.float64, requires_grad=True); b6 = torch.tensor(b, dtype=torch.float64, requires_grad=True)
    w6 = [W6[:, :4 * H], U6[0], b6[:4 * H], W6[:, 4 * H:], U6[1], b6[4 * H:]]
    mf = mb = None
    if masks is not None:
        m = torch.tensor(masks, dtype=torch.float64); mf, mb = m[:4], m[4:]
    ry = lstm_ref.bidirectional_lstm(x6, w6, mf, mb); (ry * torch.tensor(dy, dtype=torch.float64)).sum().backward()
    for name, g, r in (("y", y.detach(), ry.detach()), ("dx", xt.grad, x6.grad), ("dW", Wt.grad, W6.grad), ("dU", Ut.grad, U6.grad), ("db", bt.grad, b6.grad)):
        e = (g.cpu().double() - r).abs()
        print(masked, name, "maxerr %.3e" % float(e.max()), "ref max %.3e" % float(r.abs().max()), "n>1e-3*max:", int((e > 1e-3 * r.abs().max()).sum()), "of", e.numel())
