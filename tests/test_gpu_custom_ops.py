"""torch.library surface (torch.ops.mgr_b200.*): schema / fake-kernel / autograd registration checked with
torch.library.opcheck, and the ops agree with the module-level API they wrap."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ctc_inputs(cuda, B=6, T=40, C=22, L=7, seed=0):
    g = torch.Generator().manual_seed(seed)
    y = torch.softmax(torch.randn(B, T, C, generator=g), -1).to(cuda)
    labels = -torch.ones(B, L)
    ll = torch.zeros(B, 1, dtype=torch.int64)
    rng = np.random.default_rng(seed)
    for b in range(B):
        n = int(rng.integers(1, L + 1))
        labels[b, :n] = torch.tensor(rng.integers(0, C - 1, size=n)).float()
        ll[b, 0] = n
    il = torch.full((B, 1), T - 2, dtype=torch.int64)
    return y, labels.to(cuda), ll.to(cuda), il.to(cuda)


def test_ctc_loss_op_matches_ctc_lambda_func(cuda):
    import mgr_b200 as mgr
    from mgr_b200 import custom_ops  # noqa: F401  (registers the ops)
    y, labels, ll, il = _ctc_inputs(cuda)
    y1 = y.clone().requires_grad_(True)
    loss1 = mgr.ctc_lambda_func([y1, labels, il, ll])
    loss1.sum().backward()
    y2 = y.clone().requires_grad_(True)
    loss2, _, status = torch.ops.mgr_b200.ctc_loss(y2, labels.to(torch.int32), ll, il, False, 2, 1e-8)
    (loss2 * torch.arange(1, 7, device=cuda).float().reshape(-1, 1)).sum().backward()
    assert torch.equal(loss1, loss2) and int(status.abs().sum()) == 0
    assert torch.allclose(y2.grad, y1.grad * torch.arange(1, 7, device=cuda).float().reshape(-1, 1, 1), rtol=1e-6, atol=0)


def test_opcheck_ctc_and_decoders(cuda):
    from mgr_b200 import custom_ops  # noqa: F401
    y, labels, ll, il = _ctc_inputs(cuda, seed=3)
    yg = y.clone().requires_grad_(True)
    tests = ("test_schema", "test_faketensor", "test_autograd_registration")
    torch.library.opcheck(torch.ops.mgr_b200.ctc_loss.default, (yg, labels.to(torch.int32), ll, il, False, 2, 1e-8),
                          test_utils=tests)
    torch.library.opcheck(torch.ops.mgr_b200.ctc_bestpath.default, (y, 0.5, 2), test_utils=tests)
    torch.library.opcheck(torch.ops.mgr_b200.ctc_greedy.default, (y, None, 1e-8), test_utils=tests)
    torch.library.opcheck(torch.ops.mgr_b200.ctc_beam.default, (y, None, 8, 1, 1e-8), test_utils=tests)
    ids, lens = torch.ops.mgr_b200.ctc_bestpath(y, 0.5, 2)
    from mgr_b200 import ops
    ids2, lens2 = ops.bestpath_ref(y, 0.5, 2)
    assert torch.equal(lens, lens2)
    for n in range(y.shape[0]):
        assert torch.equal(ids[n, :lens[n]], ids2[n, :lens2[n]])


@pytest.mark.parametrize("masked", [False, True])
def test_blstm_op_matches_layer(cuda, masked):
    import mgr_b200 as mgr
    from mgr_b200 import custom_ops
    B, T, F, H = 3, 20, 24, 16
    layer = mgr.BidirectionalLSTM(F, H, dropout=0.5, seed=5).to(cuda)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, T, F, generator=g).to(cuda)
    masks = ((torch.rand(8, B, F, generator=g) > 0.5).float() * 2).to(cuda) if masked else None
    dy = torch.randn(B, T, 2 * H, generator=g).to(cuda)
    x1 = x.clone().requires_grad_(True)
    y1 = layer(x1, masks)
    g1 = torch.autograd.grad(y1, [x1, layer.kernel, layer.recurrent_kernel, layer.bias], grad_outputs=dy)
    x2 = x.clone().requires_grad_(True)
    y2 = custom_ops.blstm(x2, layer.kernel, layer.recurrent_kernel, layer.bias, masks)
    g2 = torch.autograd.grad(y2, [x2, layer.kernel, layer.recurrent_kernel, layer.bias], grad_outputs=dy)
    assert torch.equal(y1, y2)
    for a, b in zip(g1, g2):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    with torch.no_grad():
        y3 = custom_ops.blstm(x, layer.kernel, layer.recurrent_kernel, layer.bias, masks)
    assert torch.equal(y3, y1)
    torch.library.opcheck(torch.ops.mgr_b200.blstm_forward.default,
                          (x2.detach().requires_grad_(True), layer.kernel, layer.recurrent_kernel, layer.bias, masks, 0.0, True),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
