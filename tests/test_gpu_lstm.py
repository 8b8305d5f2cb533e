"""BidirectionalLSTM on the CUDA recurrence + tcgen05 projection vs the torch-CPU restatement of
Keras semantics (forward and all gradients), with and without injected input-dropout masks."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(x, W, U, b, masks, dy):
    from oracle import lstm_ref
    H = U.shape[1]
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    Wt = torch.tensor(W, dtype=torch.float64, requires_grad=True)
    Ut = torch.tensor(U, dtype=torch.float64, requires_grad=True)
    bt = torch.tensor(b, dtype=torch.float64, requires_grad=True)
    w6 = [Wt[:, :4 * H], Ut[0], bt[:4 * H], Wt[:, 4 * H:], Ut[1], bt[4 * H:]]
    mf = mb = None
    if masks is not None:
        m = torch.tensor(masks, dtype=torch.float64)
        mf, mb = m[:4], m[4:]
    y = lstm_ref.bidirectional_lstm(xt, w6, mf, mb)
    (y * torch.tensor(dy, dtype=torch.float64)).sum().backward()
    return y.detach().numpy(), xt.grad.numpy(), Wt.grad.numpy(), Ut.grad.numpy(), bt.grad.numpy()


@pytest.mark.parametrize("B,T,F,H,masked", [(4, 6, 8, 8, False), (5, 17, 20, 12, False), (3, 9, 39, 20, True),
                                            (16, 12, 64, 100, False), (7, 10, 24, 300, False),
                                            (6, 8, 40, 500, True), (32, 5, 16, 36, False)])
def test_blstm_forward_backward(cuda, B, T, F, H, masked):
    import mgr_b200 as mgr
    rng = np.random.default_rng(B * 100 + T + H)
    x = rng.standard_normal((B, T, F)).astype(np.float32)
    W = rng.uniform(-0.3, 0.3, size=(F, 8 * H)).astype(np.float32)
    U = (rng.standard_normal((2, H, 4 * H)) / np.sqrt(H)).astype(np.float32)
    b = (rng.standard_normal(8 * H) * 0.2).astype(np.float32)
    dy = rng.standard_normal((B, T, 2 * H)).astype(np.float32)
    masks = ((rng.random((8, B, F)) > 0.5) / 0.5).astype(np.float32) if masked else None
    xt = torch.tensor(x, device=cuda, requires_grad=True)
    Wt = torch.tensor(W, device=cuda, requires_grad=True)
    Ut = torch.tensor(U, device=cuda, requires_grad=True)
    bt = torch.tensor(b, device=cuda, requires_grad=True)
    mt = None if masks is None else torch.tensor(masks, device=cuda)
    y = mgr.blstm(xt, Wt, Ut, bt, mt)
    y.backward(torch.tensor(dy, device=cuda))
    torch.cuda.synchronize()
    ry, rdx, rdW, rdU, rdb = _ref(x, W, U, b, masks, dy)
    tol = lambda r: 1e-3 * max(1.0, np.abs(r).max())
    assert np.abs(y.detach().cpu().numpy() - ry).max() <= 2e-4
    assert np.abs(xt.grad.cpu().numpy() - rdx).max() <= tol(rdx)
    assert np.abs(Wt.grad.cpu().numpy() - rdW).max() <= tol(rdW)
    assert np.abs(Ut.grad.cpu().numpy() - rdU).max() <= tol(rdU)
    assert np.abs(bt.grad.cpu().numpy() - rdb).max() <= tol(rdb)


def test_keras_weight_layout_roundtrip(cuda):
    import mgr_b200 as mgr
    from oracle import lstm_ref
    rng = np.random.default_rng(0)
    F, H = 10, 6
    layer = mgr.BidirectionalLSTM(F, H).to(cuda)
    w6 = lstm_ref.init_blstm_weights(rng, F, H)
    layer.set_weights(w6)
    for a, b in zip(layer.get_weights(), w6):
        assert np.array_equal(a, b)
    x = rng.standard_normal((2, 7, F)).astype(np.float32)
    y = layer(torch.tensor(x, device=cuda)).detach().cpu().numpy()
    ref = lstm_ref.bidirectional_lstm(torch.tensor(x, dtype=torch.float64), [torch.tensor(a, dtype=torch.float64) for a in w6]).numpy()
    assert np.abs(y - ref).max() < 2e-4
    layer.forward_layer.trainable = False
    layer.backward_layer.trainable = False
    assert not any(p.requires_grad for p in layer.parameters())


def test_long_sequence_stability(cuda):
    """T=400 recurrence (config-1 length): fp32 CUDA path stays within 1e-3 of the fp64 oracle."""
    import mgr_b200 as mgr
    from oracle import lstm_ref
    rng = np.random.default_rng(9)
    B, T, F, H = 2, 400, 39, 64
    w6 = lstm_ref.init_blstm_weights(rng, F, H)
    layer = mgr.BidirectionalLSTM(F, H).to(cuda)
    layer.set_weights(w6)
    x = rng.standard_normal((B, T, F)).astype(np.float32)
    y = layer(torch.tensor(x, device=cuda)).detach().cpu().numpy()
    ref = lstm_ref.bidirectional_lstm(torch.tensor(x, dtype=torch.float64), [torch.tensor(a, dtype=torch.float64) for a in w6]).numpy()
    assert np.abs(y - ref).max() < 1e-3
