"""torch.library registration of the hot path (custom_ops.py) without a GPU: schemas, fake (meta) kernels under
FakeTensorMode, and the CUDA-only dispatch (a CPU tensor must not find a kernel -- there is no CPU fallback)."""
import pytest
import torch


def test_schemas_registered():
    import mgr_b200  # noqa: F401  (imports custom_ops)
    ops = torch.ops.mgr_b200
    for name in ("ctc_loss", "blstm_forward", "blstm_backward", "ctc_bestpath", "ctc_greedy", "ctc_beam"):
        assert hasattr(ops, name), name
    s = str(ops.ctc_loss.default._schema)
    assert "Tensor x" in s and "bool input_is_logits" in s and "-> (Tensor, Tensor, Tensor)" in s
    s = str(ops.blstm_forward.default._schema)
    assert "Tensor? masks" in s and "float mask_scale" in s and "bool keep" in s


def test_fake_kernels_shapes():
    import mgr_b200  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    B, T, C, F, H = 4, 30, 22, 39, 16
    with FakeTensorMode():
        y = torch.empty(B, T, C, device="cuda")
        lab = torch.empty(B, 5, dtype=torch.int32, device="cuda")
        ln = torch.empty(B, 1, dtype=torch.int64, device="cuda")
        loss, grad, status = torch.ops.mgr_b200.ctc_loss(y, lab, ln, ln, False, 2, 1e-8)
        assert loss.shape == (B, 1) and grad.shape == (B, T, C) and status.shape == (B,) and status.dtype == torch.int32
        x = torch.empty(B, T, F, device="cuda")
        W, U, b = torch.empty(F, 8 * H, device="cuda"), torch.empty(2, H, 4 * H, device="cuda"), torch.empty(8 * H, device="cuda")
        yb, gates, cell = torch.ops.mgr_b200.blstm_forward(x, W, U, b, None, 0.0, True)
        assert yb.shape == (B, T, 2 * H) and gates.shape == (B * T, 8 * H) and cell.shape == (B, T, 2 * H)
        yb2, g2, c2 = torch.ops.mgr_b200.blstm_forward(x, W, U, b, None, 0.0, False)
        assert yb2.shape == (B, T, 2 * H) and g2.numel() == 0 and c2.numel() == 0
        dx, dW, dU, db = torch.ops.mgr_b200.blstm_backward(x, W, U, None, gates, cell, yb, yb, 0.0, True)
        assert dx.shape == x.shape and dW.shape == W.shape and dU.shape == U.shape and db.shape == b.shape
        ids, lens = torch.ops.mgr_b200.ctc_bestpath(y, 0.5, 2)
        assert ids.shape == (B, T) and lens.shape == (B,) and ids.dtype == torch.int32
        ids, lens, logp = torch.ops.mgr_b200.ctc_beam(y, None, 100, 1, 1e-8)
        assert ids.shape == (B, 1, T) and lens.shape == (B, 1) and logp.shape == (B, 1)


def test_no_cpu_kernel():
    import mgr_b200  # noqa: F401
    y = torch.rand(2, 8, 5)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.mgr_b200.ctc_bestpath(y, 0.5, 2)
