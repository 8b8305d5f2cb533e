"""tcgen05 bf16x3 GEMM vs fp64 NumPy and vs the SIMT cross-check kernel."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(ops, A, Bm, bias, passes=3):
    a_hi, a_lo = ops.split_bf16(A)
    b_hi, b_lo = ops.split_bf16(Bm)
    M, N = A.shape[0], Bm.shape[0]
    return ops.gemm_nt(a_hi, a_lo, b_hi, b_lo, M, N, a_hi.shape[1], bias=bias, passes=passes)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 128), (300, 72, 40), (1000, 800, 1600),
                                   (77, 4000, 39), (128, 16, 8), (2048, 2400, 600), (1600, 100, 4096)])
def test_bf16x3_matches_fp64(cuda, M, N, K):
    from mgr_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(cuda)
    Bm = (torch.randn(N, K, generator=g) * 0.05).to(cuda)
    bias = torch.randn(N, generator=g).to(cuda)
    C = _gemm(ops, A, Bm, bias)
    torch.cuda.synchronize()
    ref = A.double() @ Bm.double().T + bias.double()
    err = (C.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-5 * scale, (err, scale)
    simt = ops.gemm_simt(A, Bm, bias)
    assert (simt.double() - ref).abs().max().item() <= 1e-5 * scale


def test_single_pass_bf16_is_coarser_but_sane(cuda):
    from mgr_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(1)
    A = torch.randn(256, 512, generator=g).to(cuda)
    Bm = torch.randn(192, 512, generator=g).to(cuda)
    C1 = _gemm(ops, A, Bm, None, passes=1)
    ref = A.double() @ Bm.double().T
    rel = ((C1.double() - ref).abs().max() / ref.abs().max()).item()
    assert 1e-5 < rel < 2e-2


def test_transposed_split_and_accumulate(cuda):
    """dW = X^T dP path: transposed splits, split-K accumulation, column-offset output."""
    from mgr_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(2)
    R, F, N = 4096, 200, 96
    X = torch.randn(R, F, generator=g).to(cuda)
    dP = torch.randn(R, N, generator=g).to(cuda)
    xt_hi, xt_lo = ops.split_bf16(X, transpose=True)
    pt_hi, pt_lo = ops.split_bf16(dP, transpose=True)
    out = torch.zeros(F, 2 * N, device=cuda)
    ops.gemm_nt(xt_hi, xt_lo, pt_hi, pt_lo, F, N, xt_hi.shape[1], out=out, ldc=2 * N, out_col_offset=N)
    ref = X.double().T @ dP.double()
    assert (out[:, N:].double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    assert out[:, :N].abs().max().item() == 0


def test_row_shift_split(cuda):
    from mgr_b200 import ops
    B, T, H = 3, 5, 8
    y = torch.arange(B * T * 2 * H, dtype=torch.float32, device=cuda).reshape(B * T, 2 * H) / 64
    hi, lo = ops.split_bf16(y, rows_per_seq=T, transpose=True, row_shift=-1, ncols=H, col_offset=H)
    got = (hi.float() + lo.float())[:, :B * T].T.reshape(B, T, H)
    ref = torch.zeros(B, T, H, device=cuda)
    ref[:, 1:] = y.reshape(B, T, 2 * H)[:, :-1, H:]
    assert torch.allclose(got, ref, atol=1e-6)
