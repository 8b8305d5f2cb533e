"""Parity of the CUDA CTC kernel (through the C ABI) against the oracle.
Tolerances (BASELINE.json north_star): loss 1e-4 relative, gradients 1e-3 (fp32)."""
import numpy as np
import pytest
import torch

from helpers import random_probs, random_labels

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-4
GRAD_TOL = 1e-3


def _run(mgr, dev, p, labels, il, ll, logits=None):
    if logits is None:
        y = torch.tensor(p, device=dev, requires_grad=True)
        loss = mgr.ctc_lambda_func([y, torch.tensor(labels), torch.tensor(il), torch.tensor(ll)])
    else:
        y = torch.tensor(logits, device=dev, requires_grad=True)
        loss = mgr.softmax_ctc(y, torch.tensor(labels), torch.tensor(il), torch.tensor(ll))
    loss.mean().backward()
    return loss.detach().cpu().numpy().astype(np.float64), y.grad.detach().cpu().numpy().astype(np.float64)


@pytest.mark.parametrize("B,T,C,Lmax", [(4, 12, 5, 3), (8, 50, 22, 10), (3, 33, 44, 40), (5, 130, 22, 35),
                                        (2, 70, 6, 33), (2, 300, 22, 150)])
def test_loss_and_grad_match_oracle(cuda, B, T, C, Lmax):
    import mgr_b200 as mgr
    from oracle import ctc_ref
    rng = np.random.default_rng(B * 1000 + T)
    p, _ = random_probs(rng, B, T, C)
    il = rng.integers((T - 2) // 2 + 1, T - 1, size=(B, 1))
    il[0, 0] = T - 2
    labels, ll = random_labels(rng, B, Lmax, C, T_avail=il[:, 0])
    loss, grad = _run(mgr, cuda, p, labels, il, ll)
    ref_loss, ref_g = ctc_ref.ctc_lambda_func((p, labels, il, ll), want_grad=True)
    ref_g = ref_g / B
    assert np.abs(loss - ref_loss).max() <= LOSS_RTOL * np.abs(ref_loss).max()
    # gradient wrt probabilities: compare relative to the row scale (1/(p+eps) can be large)
    scale = np.abs(ref_g).max(axis=2, keepdims=True) + 1e-12
    assert (np.abs(grad - ref_g) / scale).max() <= GRAD_TOL
    assert np.all(grad[:, :2] == 0)
    for b in range(B):
        assert np.all(grad[b, 2 + il[b, 0]:] == 0)


@pytest.mark.parametrize("B,T,C,Lmax", [(4, 40, 22, 8), (2, 200, 22, 35), (3, 64, 44, 20)])
def test_fused_logit_mode(cuda, B, T, C, Lmax):
    import mgr_b200 as mgr
    from oracle import ctc_ref
    rng = np.random.default_rng(7 + T)
    _, a = random_probs(rng, B, T, C)
    il = np.full((B, 1), T - 2)
    labels, ll = random_labels(rng, B, Lmax, C, T_avail=il[:, 0])
    loss, grad = _run(mgr, cuda, None, labels, il, ll, logits=a)
    ref_loss, ref_g = ctc_ref.softmax_ctc_grad_logits(a, labels, il, ll)
    assert np.abs(loss - ref_loss).max() <= LOSS_RTOL * np.abs(ref_loss).max()
    assert np.abs(grad - ref_g).max() <= GRAD_TOL * np.abs(ref_g).max()


@pytest.mark.parametrize("impl", ["v5", "v4"])
@pytest.mark.parametrize("Tn", [1, 2, 3, 7, 31, 32, 33, 63, 64, 65, 66, 97, 128, 129])
def test_chunk_boundaries_and_both_kernels(cuda, monkeypatch, impl, Tn):
    """Sequence lengths around the 32-frame chunk, the meeting row t* = Tn/2 and the 4-row lattice prefetch
    window; the round-1 kernel (GR_CTC_IMPL=v4) stays as a cross-check."""
    import mgr_b200 as mgr
    from oracle import ctc_ref
    monkeypatch.setenv("GR_CTC_IMPL", impl)
    B, C, Lmax, T = 3, 22, 12, 140
    rng = np.random.default_rng(100 + Tn)
    p, a = random_probs(rng, B, T, C)
    il = np.full((B, 1), Tn)
    il[1, 0] = max(1, Tn - 1)
    labels, ll = random_labels(rng, B, Lmax, C, T_avail=np.maximum(il[:, 0], 2))
    ll = np.minimum(ll, il)
    labels[0, 1:] = -1
    ll[0, 0] = 1
    loss, grad = _run(mgr, cuda, p, labels, il, ll)
    ref_loss, ref_g = ctc_ref.ctc_lambda_func((p, labels, il, ll), want_grad=True)
    fin = np.isfinite(ref_loss)
    assert np.array_equal(np.isfinite(loss), fin)
    assert np.abs(loss[fin] - ref_loss[fin]).max() <= LOSS_RTOL * np.abs(ref_loss[fin]).max()
    ref_g = ref_g / B
    scale = np.abs(ref_g).max(axis=2, keepdims=True) + 1e-12
    assert (np.abs(grad - ref_g) / scale).max() <= GRAD_TOL
    loss2, grad2 = _run(mgr, cuda, None, labels, il, ll, logits=a)
    ref_loss2, ref_g2 = ctc_ref.softmax_ctc_grad_logits(a, labels, il, ll)
    assert np.abs(loss2[fin] - ref_loss2[fin]).max() <= LOSS_RTOL * np.abs(ref_loss2[fin]).max()
    assert np.abs(grad2 - ref_g2).max() <= GRAD_TOL * np.abs(ref_g2).max()


@pytest.mark.parametrize("tc", ["16", "8"])
def test_small_chunk_instantiations(cuda, monkeypatch, tc):
    """The chunk size is chosen from the batch size (shared memory for a single wave); small test batches always
    get 32 frames, so force the other two template instantiations (GR_CTC_TC).  (First run on a GPU in round 2,
    also under compute-sanitizer memcheck: profiles/r02_sanitizer_ctc_small_chunks.txt.)"""
    import mgr_b200 as mgr
    from oracle import ctc_ref
    monkeypatch.setenv("GR_CTC_TC", tc)
    B, T, C, Lmax = 3, 77, 22, 12
    rng = np.random.default_rng(int(tc))
    p, a = random_probs(rng, B, T, C)
    il = np.array([[T - 2], [41], [17]])
    labels, ll = random_labels(rng, B, Lmax, C, T_avail=il[:, 0])
    loss, grad = _run(mgr, cuda, p, labels, il, ll)
    ref_loss, ref_g = ctc_ref.ctc_lambda_func((p, labels, il, ll), want_grad=True)
    assert np.abs(loss - ref_loss).max() <= LOSS_RTOL * np.abs(ref_loss).max()
    ref_g = ref_g / B
    scale = np.abs(ref_g).max(axis=2, keepdims=True) + 1e-12
    assert (np.abs(grad - ref_g) / scale).max() <= GRAD_TOL
    loss2, grad2 = _run(mgr, cuda, None, labels, il, ll, logits=a)
    ref_loss2, ref_g2 = ctc_ref.softmax_ctc_grad_logits(a, labels, il, ll)
    assert np.abs(loss2 - ref_loss2).max() <= LOSS_RTOL * np.abs(ref_loss2).max()
    assert np.abs(grad2 - ref_g2).max() <= GRAD_TOL * np.abs(ref_g2).max()


def test_edge_cases(cuda):
    import mgr_b200 as mgr
    from oracle import ctc_ref
    C = 5
    rng = np.random.default_rng(3)
    p, _ = random_probs(rng, 5, 9, C)
    # blank-only label -> empty target; T'=1; repeated labels just feasible; B=1 batch (Keras squeeze bug n/a)
    labels = np.array([[4, -1, -1], [1, -1, -1], [2, 2, -1], [0, 1, 0], [3, 4, -1]], dtype=np.float32)
    ll = np.array([[1], [1], [2], [3], [2]])
    il = np.array([[7], [1], [3], [5], [7]])
    loss, grad = _run(mgr, cuda, p, labels, il, ll)
    ref_loss, ref_g = ctc_ref.ctc_lambda_func((p, labels, il, ll), want_grad=True)
    assert np.abs(loss - ref_loss).max() <= LOSS_RTOL * np.abs(ref_loss).max()
    scale = np.abs(ref_g).max(axis=2, keepdims=True) + 1e-12
    assert (np.abs(grad - ref_g / 5) / scale).max() <= GRAD_TOL
    l1, _ = _run(mgr, cuda, p[:1], labels[:1], il[:1], ll[:1])
    assert abs(l1[0, 0] - ref_loss[0, 0]) <= LOSS_RTOL * abs(ref_loss[0, 0])


def test_no_valid_path_gives_inf_like_tf(cuda):
    import mgr_b200 as mgr
    p = torch.full((1, 4, 4), 0.25, device=cuda)
    loss = mgr.ctc_lambda_func([p, torch.tensor([[1.0, 1.0]]), torch.tensor([[2]]), torch.tensor([[2]])])
    assert torch.isinf(loss).all()


@pytest.mark.parametrize("labels,ll,il", [([[0.0]], [[0]], [[6]]), ([[3.0, 1.0]], [[2]], [[6]]),
                                           ([[0.0] * 7], [[7]], [[6]]), ([[0.0]], [[1]], [[9]])])
def test_tf_error_behaviour(cuda, labels, ll, il):
    import mgr_b200 as mgr
    p = torch.full((1, 8, 4), 0.25, device=cuda)
    with pytest.raises(mgr.InvalidArgumentError):
        mgr.ctc_lambda_func([p, torch.tensor(labels), torch.tensor(il), torch.tensor(ll)])


def test_full_size_properties(cuda):
    """BASELINE config 4 size (B=1024, T=1000, L<=40, C=22): size-independent properties --
    grad wrt logits sums to zero over classes, frames 0,1 are zero, and a spot-checked subset of
    sequences matches the oracle."""
    import mgr_b200 as mgr
    from oracle import ctc_ref
    B, T, C, Lmax = 1024, 1002, 22, 40
    g = torch.Generator(device="cpu").manual_seed(3001)
    a = (torch.randn(B, T, C, generator=g) * 2).to(cuda)
    rng = np.random.default_rng(3002)
    labels, ll = random_labels(rng, B, Lmax, C)
    il = np.full((B, 1), T - 2)
    y = a.clone().requires_grad_(True)
    loss = mgr.softmax_ctc(y, torch.tensor(labels), torch.tensor(il), torch.tensor(ll))
    loss.sum().backward()
    gsum = y.grad.sum(dim=2).abs().max().item()
    assert gsum < 1e-4
    assert torch.all(y.grad[:, :2] == 0)
    assert torch.isfinite(loss).all()
    idx = sorted(set([0, 511, 1023]) | set(np.random.default_rng(7).choice(B, 64, replace=False).tolist()))
    ref_loss, ref_g = ctc_ref.softmax_ctc_grad_logits(a[idx].cpu().numpy(), labels[idx], il[idx], ll[idx],
                                                      upstream=np.ones(len(idx)))
    got = loss[idx].detach().cpu().numpy()
    assert np.abs(got - ref_loss).max() <= LOSS_RTOL * np.abs(ref_loss).max()
    gg = y.grad[idx].cpu().numpy()
    assert np.abs(gg - ref_g).max() <= GRAD_TOL * np.abs(ref_g).max()


@pytest.mark.parametrize("B,T,C,Lmax,ragged", [(1024, 402, 44, 40, False), (512, 602, 44, 150, True), (700, 302, 22, 150, True)])
def test_large_batch_chunk_selection(cuda, B, T, C, Lmax, ragged):
    """Shapes at which `launch_ctc` picks the 16- / 8-frame chunk instantiations BY ITSELF (one wave of CTAs no longer
    fits with 32-frame chunks): BASELINE config 4's "also C=44" at B=1024, and the speech net's Lmax=150
    (audio_network/data_generator.py absolute_max_sequence_len) at large batch.  64 random sequences vs the oracle."""
    import mgr_b200 as mgr
    from oracle import ctc_ref
    g = torch.Generator(device="cpu").manual_seed(B + C)
    a = (torch.randn(B, T, C, generator=g) * 2).to(cuda)
    rng = np.random.default_rng(B * 7 + Lmax)
    il = rng.integers((T - 2) // 2, T - 1, size=(B, 1)) if ragged else np.full((B, 1), T - 2)
    labels, ll = random_labels(rng, B, Lmax, C, T_avail=il[:, 0])
    y = a.clone().requires_grad_(True)
    loss = mgr.softmax_ctc(y, torch.tensor(labels), torch.tensor(il), torch.tensor(ll))
    loss.sum().backward()
    fin = torch.isfinite(loss).cpu().numpy().ravel()
    assert y.grad.sum(dim=2).abs().max().item() < 1e-4
    idx = np.random.default_rng(11).choice(B, 64, replace=False)
    ref_loss, ref_g = ctc_ref.softmax_ctc_grad_logits(a[idx].cpu().numpy(), labels[idx], il[idx], ll[idx],
                                                      upstream=np.ones(len(idx)))
    got = loss[idx].detach().cpu().numpy()
    f = np.isfinite(ref_loss).ravel()
    assert np.array_equal(f, fin[idx])
    assert np.abs(got.ravel()[f] - ref_loss.ravel()[f]).max() <= LOSS_RTOL * np.abs(ref_loss.ravel()[f]).max()
    gg = y.grad[idx].cpu().numpy()
    assert np.abs(gg[f] - ref_g[f]).max() <= GRAD_TOL * np.abs(ref_g[f]).max()
