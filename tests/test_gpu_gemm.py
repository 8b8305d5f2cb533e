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


@pytest.mark.parametrize("BT,T,F,H,nvar", [(96, 12, 64, 32, 8), (200, 25, 39, 20, 8), (384, 48, 1600, 100, 8),
                                           (300, 30, 20, 300, 1), (256, 16, 1000, 500, 8),
                                           (1000, 200, 1600, 100, 8), (768, 128, 40, 500, 8), (1300, 650, 600, 300, 8),
                                           (900, 300, 1000, 500, 1), (1000, 200, 40, 300, 8), (333, 111, 24, 500, 8),
                                           (520, 130, 64, 150, 1)])
@pytest.mark.parametrize("mask_scale", [0.0, 2.0])
def test_fused_prologue_projection(cuda, BT, T, F, H, nvar, mask_scale):
    """gr_gemm_a32_f32, transA=0: P[:, v*H:(v+1)*H] = (X o mask_v) W_v + b with X read as fp32.
    mask_scale = 2: the same masks declared as dropout masks {0, 2} (gr_gemm_a32_dropout_f32: one split per tile
    shared by four variants, scale in the epilogue) -- same tolerance against the fp64 product."""
    if mask_scale and nvar != 8:
        pytest.skip("no masks")
    from mgr_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(BT + F)
    X = torch.randn(BT, F, generator=g).to(cuda)
    Nv = H if nvar == 8 else 8 * H
    W = (torch.randn(F, nvar * Nv, generator=g) * 0.1).to(cuda)
    bias = torch.randn(nvar * Nv, generator=g).to(cuda)
    masks = ((torch.rand(nvar, BT // T, F, generator=g) > 0.5).float() * 2).to(cuda) if nvar == 8 else None
    w_hi, w_lo = ops.split_bf16(W, transpose=True)
    out = torch.empty(BT, nvar * Nv, device=cuda)
    ops.gemm_a32(X, w_hi, w_lo, BT, Nv, F, out, nvar * Nv, nvar=nvar, mask=masks, rows_per_seq=T, bias=bias,
                 mask_scale=mask_scale)
    torch.cuda.synchronize()
    Xd, Wd = X.double(), W.double()
    if masks is None:
        ref = Xd @ Wd + bias.double()
    else:
        ref = torch.cat([(Xd.reshape(-1, T, F) * masks[v].double()[:, None, :]).reshape(BT, F) @ Wd[:, v * Nv:(v + 1) * Nv]
                         for v in range(nvar)], 1) + bias.double()
    assert (out.double() - ref).abs().max().item() <= 3e-5 * ref.abs().max().item()


@pytest.mark.parametrize("BT,T,F,H,masked", [(256, 32, 64, 32, True), (1024, 64, 1600, 100, True), (600, 50, 39, 24, False),
                                             (4096, 128, 200, 100, True), (3000, 1000, 1600, 100, True),
                                             (2000, 100, 77, 36, True), (1500, 500, 600, 300, False),
                                             (5000, 1000, 1600, 100, True), (4224, 66, 200, 100, True)])
@pytest.mark.parametrize("mask_scale", [0.0, 2.0])
def test_fused_prologue_weight_gradient(cuda, BT, T, F, H, masked, mask_scale):
    """transA=1: dW_v = (X o mask_v)^T dP_v (split-K, atomics) and the shifted dU contraction; mask_scale = 2: the
    masks declared as dropout masks (shared split, keep bits; T = 66 / 1000 put sequence ends inside 8-row chunks)."""
    if mask_scale and not masked:
        pytest.skip("no masks")
    from mgr_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(BT + F + 1)
    X = torch.randn(BT, F, generator=g).to(cuda)
    dP = torch.randn(BT, 8 * H, generator=g).to(cuda)
    masks = ((torch.rand(8, BT // T, F, generator=g) > 0.5).float() * 2).to(cuda) if masked else None
    pt_hi, pt_lo = ops.split_bf16(dP, transpose=True)
    dW = torch.empty(F, 8 * H, device=cuda)
    if masked:
        ops.gemm_a32(X, pt_hi, pt_lo, F, H, BT, dW, 8 * H, nvar=8, mask=masks, rows_per_seq=T, transA=True,
                     mask_scale=mask_scale)
        ref = torch.cat([(X.double().reshape(-1, T, F) * masks[v].double()[:, None, :]).reshape(BT, F).T
                         @ dP.double()[:, v * H:(v + 1) * H] for v in range(8)], 1)
    else:
        ops.gemm_a32(X, pt_hi, pt_lo, F, 8 * H, BT, dW, 8 * H, transA=True, rows_per_seq=T)
        ref = X.double().T @ dP.double()
    torch.cuda.synchronize()
    assert (dW.double() - ref).abs().max().item() <= 3e-5 * ref.abs().max().item()
    # dU_d = Hprev^T dP_d with the time shift inside each sequence
    Y = torch.randn(BT, 2 * H, generator=g).to(cuda)
    for d, shift in ((0, -1), (1, 1)):
        dU = torch.empty(H, 4 * H, device=cuda)
        ops.gemm_a32(Y, pt_hi[d * 4 * H:(d + 1) * 4 * H], pt_lo[d * 4 * H:(d + 1) * 4 * H], H, 4 * H, BT, dU, 4 * H,
                     transA=True, rows_per_seq=T, row_shift=shift, a_col_offset=d * H)
        Y3 = Y.double().reshape(-1, T, 2 * H)[:, :, d * H:(d + 1) * H]
        Hp = torch.zeros_like(Y3)
        if shift == -1:
            Hp[:, 1:] = Y3[:, :-1]
        else:
            Hp[:, :-1] = Y3[:, 1:]
        refU = Hp.reshape(BT, H).T @ dP.double()[:, d * 4 * H:(d + 1) * 4 * H]
        torch.cuda.synchronize()
        assert (dU.double() - refU).abs().max().item() <= 3e-5 * refU.abs().max().item()


@pytest.mark.parametrize("nvg", ["1", "4"])
def test_projection_variant_grouping(cuda, monkeypatch, nvg):
    """Both tile schedules of the fused projection (one variant per CTA / four variants sharing one
    loaded A tile, GR_A32_NVG) give the same parity, row-major and transposed."""
    monkeypatch.setenv("GR_A32_NVG", nvg)
    test_fused_prologue_projection(cuda, 384, 48, 1600, 100, 8, 0.0)
    test_fused_prologue_projection(cuda, 1000, 200, 1600, 100, 8, 0.0)
    test_fused_prologue_weight_gradient(cuda, 384, 96, 1600, 100, True, 0.0)


def test_projection_epilogue_variants(cuda, monkeypatch):
    """The TMA-store epilogue and the st.global cross-check epilogue (GR_A32_EPI=stg) agree with fp64."""
    monkeypatch.setenv("GR_A32_EPI", "stg")
    test_fused_prologue_projection(cuda, 256, 16, 1000, 500, 8, 0.0)
    test_fused_prologue_projection(cuda, 768, 128, 40, 500, 8, 0.0)
