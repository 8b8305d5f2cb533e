"""Pins for oracle/lstm_ref.py: hand-computed single step, Bidirectional wiring, gradcheck."""
import numpy as np
import torch

from oracle import lstm_ref


def test_single_step_by_hand():
    rng = np.random.default_rng(0)
    F, H = 3, 2
    k, r, b = rng.standard_normal((F, 4 * H)), rng.standard_normal((H, 4 * H)), rng.standard_normal(4 * H)
    x = rng.standard_normal((1, 2, F))
    y = lstm_ref.keras_lstm(torch.tensor(x), torch.tensor(k), torch.tensor(r), torch.tensor(b)).numpy()
    hs = lambda v: np.clip(0.2 * v + 0.5, 0, 1)
    z = x[0, 0] @ k + b
    i, f, g, o = hs(z[:H]), hs(z[H:2 * H]), np.tanh(z[2 * H:3 * H]), hs(z[3 * H:])
    c = i * g
    h = o * np.tanh(c)
    assert np.allclose(y[0, 0], h)
    z = x[0, 1] @ k + b + h @ r
    i, f, g, o = hs(z[:H]), hs(z[H:2 * H]), np.tanh(z[2 * H:3 * H]), hs(z[3 * H:])
    c2 = f * c + i * g
    assert np.allclose(y[0, 1], o * np.tanh(c2))


def test_bidirectional_wiring():
    rng = np.random.default_rng(1)
    B, T, F, H = 2, 5, 3, 4
    w = [torch.tensor(a, dtype=torch.float64) for a in lstm_ref.init_blstm_weights(rng, F, H, np.float64)]
    x = torch.tensor(rng.standard_normal((B, T, F)))
    y = lstm_ref.bidirectional_lstm(x, w)
    assert y.shape == (B, T, 2 * H)
    yb = lstm_ref.keras_lstm(torch.flip(x, [1]), w[3], w[4], w[5])
    assert torch.allclose(y[:, :, H:], torch.flip(yb, [1]))
    # forward half at t=0 only sees x[0]; backward half at t=T-1 only sees x[T-1]
    x2 = x.clone()
    x2[:, 1:] += 1.0
    y2 = lstm_ref.bidirectional_lstm(x2, w)
    assert torch.allclose(y[:, 0, :H], y2[:, 0, :H]) and not torch.allclose(y[:, 0, H:], y2[:, 0, H:])


def test_gradcheck_with_masks():
    rng = np.random.default_rng(2)
    B, T, F, H = 2, 4, 3, 2
    k = torch.tensor(rng.standard_normal((F, 4 * H)) * 0.3, requires_grad=True)
    r = torch.tensor(rng.standard_normal((H, 4 * H)) * 0.3, requires_grad=True)
    b = torch.tensor(rng.standard_normal(4 * H) * 0.1, requires_grad=True)
    x = torch.tensor(rng.standard_normal((B, T, F)), requires_grad=True)
    masks = torch.tensor((rng.random((4, B, F)) > 0.4) / 0.6)
    assert torch.autograd.gradcheck(lambda *a: lstm_ref.keras_lstm(*a, go_backwards=True, masks=masks), (x, k, r, b), eps=1e-6, atol=1e-5)
