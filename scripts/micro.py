"""Micro-workloads for ncu captures: `python scripts/micro.py ctc|lstm_tc|gemm`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mgr_b200 import ops, layers
dev = torch.device("cuda:0")
what = sys.argv[1]
if what == "ctc":
    B, T, C, L = 1024, 1002, 22, 40
    g = torch.Generator().manual_seed(3001)
    probs = torch.softmax(torch.randn(B, T, C, generator=g) * 2, -1).to(dev)
    rng = np.random.default_rng(3002)
    labels = torch.tensor(rng.integers(0, C - 1, size=(B, L)), dtype=torch.int32, device=dev)
    ll = torch.full((B,), L, dtype=torch.int32, device=dev); il = torch.full((B,), T - 2, dtype=torch.int32, device=dev)
    for _ in range(3):
        ops.ctc_loss_grad(probs, labels, ll, il, False)
elif what == "lstm_tc":
    B, T, H = 256, 1000, 500
    gates = torch.randn(B * T, 8 * H, device=dev) * 0.5
    U = torch.randn(2, H, 4 * H, device=dev) / H ** 0.5
    for _ in range(2):
        ops.lstm_recurrence_fwd(gates.clone(), U, B, T, H, keep_cell=False)
elif what == "lstm_tcu":       # the default forward recurrence (lstm_tcu.cu): speech tower layer, inference
    B, T, H = 256, 1000, 500
    gates = torch.randn(B * T, 8 * H, device=dev) * 0.5
    U = torch.randn(2, H, 4 * H, device=dev) / H ** 0.5
    for _ in range(2):
        ops.lstm_recurrence_fwd(gates.clone(), U, B, T, H, keep_cell=False)
elif what == "lstm_tcu_bwd":   # tensor-core BPTT at the config-2 layer size
    B, T, H = 64, 800, 300
    gates = torch.randn(B * T, 8 * H, device=dev) * 0.5
    U = torch.randn(2, H, 4 * H, device=dev) / H ** 0.5
    dy = torch.randn(B, T, 2 * H, device=dev) * 0.1
    y, cell = ops.lstm_recurrence_fwd(gates, U, B, T, H, keep_cell=True)
    for _ in range(2):
        ops.lstm_recurrence_bwd(gates.clone(), cell, dy, U, B, T, H)
elif what == "gemm":
    M, N, K = 65536, 4000, 1000
    x = torch.randn(M, K, device=dev); W = torch.randn(K, N, device=dev) * 0.05
    a = ops.split_bf16(x); w = ops.split_bf16(W, transpose=True)
    for _ in range(2):
        ops.gemm_nt(a[0], a[1], w[0], w[1], M, N, K)
elif what == "a32":
    BT, T, F, H = 65536, 1000, 1000, 500
    x = torch.randn(BT, F, device=dev); W = torch.randn(F, 8 * H, device=dev) * 0.05; b = torch.zeros(8 * H, device=dev)
    masks = ((torch.rand(8, BT // T + 1, F, device=dev) > 0.5).float() * 2)[:, :BT // T + 1].contiguous()
    for _ in range(2):
        layers._project(x, W, b, masks, BT // T, T, H)
elif what == "a32full":       # speech layer-2 projection at the full benchmark batch
    BT, T, F, H = 256000, 1000, 1000, 500
    x = torch.randn(BT, F, device=dev); W = torch.randn(F, 8 * H, device=dev) * 0.05; b = torch.zeros(8 * H, device=dev)
    masks = ((torch.rand(8, BT // T, F, device=dev) > 0.5).float() * 2).contiguous()
    for _ in range(2):
        layers._project(x, W, b, masks, BT // T, T, H, mask_scale=2.0)
elif what == "small":         # fusion BLSTM(100) forward + backward at the benchmark batch
    B, T, H = 256, 1000, 100
    gates = torch.randn(B * T, 8 * H, device=dev) * 0.5
    U = torch.randn(2, H, 4 * H, device=dev) / H ** 0.5
    dy = torch.randn(B, T, 2 * H, device=dev)
    for _ in range(2):
        g2 = gates.clone()
        y, c = ops.lstm_recurrence_fwd(g2, U, B, T, H, keep_cell=True)
        ops.lstm_recurrence_bwd(g2, c, dy, U, B, T, H)
elif what == "a32t":
    BT, T, F, H = 65536, 1000, 1600, 100
    x = torch.randn(BT, F, device=dev); dP = torch.randn(BT, 8 * H, device=dev)
    masks = ((torch.rand(8, BT // T + 1, F, device=dev) > 0.5).float() * 2).contiguous()
    pt = ops.split_bf16(dP, transpose=True)
    dW = torch.empty(F, 8 * H, device=dev)
    for _ in range(2):
        ops.gemm_a32(x, pt[0], pt[1], F, H, BT, dW, 8 * H, nvar=8, mask=masks, rows_per_seq=T, transA=True)
torch.cuda.synchronize()
