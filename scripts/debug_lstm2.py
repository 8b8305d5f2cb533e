import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgr_b200 as mgr
from mgr_b200 import ops, layers
from oracle import lstm_ref
dev = torch.device("cuda:0")

def run(B, T, F, H, u_mode, b_mode, w_scale, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, T, F)).astype(np.float32)
    W = rng.uniform(-w_scale, w_scale, size=(F, 8 * H)).astype(np.float32)
    if u_mode == "orth":
        U = np.stack([lstm_ref.orthogonal(rng, H, 4 * H) for _ in range(2)]).astype(np.float32)
    else:
        U = (rng.standard_normal((2, H, 4 * H)) / np.sqrt(H)).astype(np.float32)
    if b_mode == "keras":
        b = np.zeros(8 * H, np.float32); b[H:2*H] = 1; b[5*H:6*H] = 1
    elif b_mode == "zero":
        b = np.zeros(8 * H, np.float32)
    else:
        b = (rng.standard_normal(8 * H) * 0.2).astype(np.float32)
    xt, Wt, Ut, bt = [torch.tensor(a, device=dev) for a in (x, W, U, b)]
    P = layers._project(xt.reshape(B * T, F), Wt, bt, None, B, T, H, 3)
    Pref = torch.tensor(x.reshape(B * T, F).astype(np.float64) @ W.astype(np.float64) + b)
    perr = float((P.cpu().double() - Pref).abs().max())
    Pc = P.clone()
    y, cell = ops.lstm_recurrence_fwd(P, Ut, B, T, H, keep_cell=True)
    torch.cuda.synchronize()
    # reference recurrence from the GPU's own P (float64 on CPU)
    Pn = Pc.cpu().double().reshape(B, T, 8 * H); Un = torch.tensor(U, dtype=torch.float64)
    yr = torch.zeros(B, T, 2 * H, dtype=torch.float64)
    hs = lambda v: torch.clamp(0.2 * v + 0.5, 0, 1)
    for d in range(2):
        h = torch.zeros(B, H, dtype=torch.float64); c = torch.zeros(B, H, dtype=torch.float64)
        for s in range(T):
            t = s if d == 0 else T - 1 - s
            z = Pn[:, t, d * 4 * H:(d + 1) * 4 * H] + h @ Un[d]
            i, f, g, o = hs(z[:, :H]), hs(z[:, H:2*H]), torch.tanh(z[:, 2*H:3*H]), hs(z[:, 3*H:])
            c = f * c + i * g; h = o * torch.tanh(c); yr[:, t, d * H:(d + 1) * H] = h
    e = (y.cpu().double() - yr).abs()
    et = e.amax(dim=(0, 2))
    print("B%d T%d F%d H%d U=%s b=%s w=%.2f | Perr %.2e | y err %.2e fwd %.2e bwd %.2e | first bad step fwd-dir: %s"
          % (B, T, F, H, u_mode, b_mode, w_scale, perr, float(e.max()), float(e[..., :H].max()), float(e[..., H:].max()),
             [round(float(v), 5) for v in e[..., :H].amax(dim=(0, 2))[:8]]))

for cfg in [(2, 7, 10, 6, "orth", "keras", 0.05), (2, 7, 10, 6, "rand", "keras", 0.05), (2, 7, 10, 6, "orth", "rand", 0.05),
            (2, 7, 10, 6, "orth", "zero", 0.05), (2, 7, 10, 6, "orth", "keras", 0.3), (4, 7, 10, 6, "orth", "keras", 0.05),
            (2, 7, 16, 8, "orth", "keras", 0.05), (4, 6, 8, 8, "rand", "rand", 0.3), (4, 40, 39, 24, "orth", "keras", 0.05)]:
    run(*cfg)
