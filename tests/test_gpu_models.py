"""Whole-topology parity (speech / skeletal / fusion) and one fusion training step vs the oracle,
at reduced widths so the CPU oracle finishes in seconds; plus the head and optimiser kernels."""
import numpy as np
import pytest
import torch

from helpers import random_labels

pytestmark = pytest.mark.gpu


def _t64(ws):
    return [torch.tensor(w, dtype=torch.float64) for w in ws]


def _small_nets(mgr, cuda, Fa=39, Fs=20, Ha=24, Hs=16, Hf=12, C=22):
    sp = mgr.UnimodalNet(Fa, Ha, 44, 0.5, (0.4, 0.5, 0.5), seed=1).to(cuda)
    sk = mgr.UnimodalNet(Fs, Hs, C, 0.5, (0.6, 0.6, 0.6), seed=2).to(cuda)
    fu = mgr.FusionNet(sp, sk, nb_classes=C, units=Hf, seed=3).to(cuda)
    return sp, sk, fu


def test_unimodal_forward_and_loss(cuda):
    import mgr_b200 as mgr
    from oracle import lstm_ref, ctc_ref
    rng = np.random.default_rng(1001)
    B, T, F, H, C = 4, 40, 39, 24, 21
    net = mgr.UnimodalNet(F, H, C, 0.5, (0.4, 0.5, 0.5), seed=5).to(cuda)
    x = rng.standard_normal((B, T, F)).astype(np.float32)
    labels, ll = random_labels(rng, B, 8, C)
    il = np.full((B, 1), T - 2)
    y_pred, logits = net(torch.tensor(x, device=cuda))
    loss = mgr.ctc_lambda_func([y_pred, torch.tensor(labels), torch.tensor(il), torch.tensor(ll)])
    p_ref, _, _ = lstm_ref.unimodal_forward(torch.tensor(x, dtype=torch.float64), _t64(net.blstm_1.get_weights()),
                                            _t64(net.blstm_2.get_weights()), _t64(net.dense.get_weights()))
    assert np.abs(y_pred.detach().cpu().numpy() - p_ref.numpy()).max() < 2e-4
    ref_loss = ctc_ref.ctc_lambda_func((p_ref.numpy(), labels, il, ll))
    assert np.abs(loss.detach().cpu().numpy() - ref_loss).max() <= 1e-4 * np.abs(ref_loss).max()


def test_fusion_training_step_matches_oracle(cuda):
    import mgr_b200 as mgr
    from oracle import lstm_ref
    rng = np.random.default_rng(2001)
    B, T, C = 4, 30, 22
    sp, sk, fu = _small_nets(mgr, cuda, C=C)
    xa = rng.standard_normal((B, T, 39)).astype(np.float32)
    xs = rng.standard_normal((B, T, 20)).astype(np.float32)
    labels, ll = random_labels(rng, B, 6, C)
    il = np.full((B, 1), T - 2)
    # injected regularisers (RNG streams cannot match TF's: SURVEY 8a a6)
    reg = {"sp": {"noise": torch.tensor(rng.standard_normal((B, T, 39)).astype(np.float32) * 0.5, device=cuda)},
           "sk": {},
           "m3": torch.tensor(((rng.random((8, B, 80)) > 0.5) / 0.5).astype(np.float32), device=cuda),
           "drop": torch.tensor(((rng.random((B, T, 24)) > 0.5) / 0.5).astype(np.float32), device=cuda)}
    loss, grads = fu.loss_and_grads(torch.tensor(xa, device=cuda), torch.tensor(xs, device=cuda), labels, il, ll, reg,
                                    check=True)
    # oracle
    fw = [w.clone().requires_grad_(True) for w in _t64(fu.blstm_3.get_weights())]
    fd = [w.clone().requires_grad_(True) for w in _t64(fu.dense.get_weights())]
    m3 = reg["m3"].cpu().double()
    p, a, _ = lstm_ref.fusion_forward(torch.tensor(xa, dtype=torch.float64), torch.tensor(xs, dtype=torch.float64),
                                      _t64(sp.blstm_1.get_weights()), _t64(sp.blstm_2.get_weights()),
                                      _t64(sk.blstm_1.get_weights()), _t64(sk.blstm_2.get_weights()), fw, fd,
                                      noise_a=reg["sp"]["noise"].cpu().double(),
                                      masks={"fu_f": m3[:4], "fu_b": m3[4:]}, drop_mask=reg["drop"].cpu().double())
    ref_loss = lstm_ref.torch_ctc_lambda(p, labels, il, ll)
    ref_loss.mean().backward()
    rl = ref_loss.detach().numpy()[:, 0]
    assert np.abs(loss.cpu().numpy() - rl).max() <= 1e-4 * np.abs(rl).max()
    H = 12
    ref_W = torch.cat([fw[0].grad, fw[3].grad], 1).numpy()
    ref_U = torch.stack([fw[1].grad, fw[4].grad]).numpy()
    ref_b = torch.cat([fw[2].grad, fw[5].grad]).numpy()
    for got, ref in zip(grads, [ref_W, ref_U, ref_b, fd[0].grad.numpy(), fd[1].grad.numpy()]):
        assert np.abs(got.cpu().numpy() - ref).max() <= 1e-3 * max(np.abs(ref).max(), 1e-6)
    # autograd surface gives the same gradients as the explicit step
    y_pred, _ = fu(torch.tensor(xa, device=cuda), torch.tensor(xs, device=cuda), reg)
    l2 = mgr.ctc_lambda_func([y_pred, torch.tensor(labels), torch.tensor(il), torch.tensor(ll)])
    l2.mean().backward()
    assert np.abs(fu.blstm_3.kernel.grad.cpu().numpy() - ref_W).max() <= 1e-3 * np.abs(ref_W).max()
    assert np.abs(fu.dense.kernel.grad.cpu().numpy() - fd[0].grad.numpy()).max() <= 1e-3 * np.abs(fd[0].grad.numpy()).max()
    assert sp.blstm_1.kernel.grad is None  # frozen towers


def test_half_batch_speech_towers_match_whole_batch(cuda, monkeypatch):
    """GR_TOWER_SPLIT=1 runs the speech tower as two half-batch towers on two streams beside the skeletal one
    (so the recurrences fit on the GPU together); the merged features must not depend on that (sequences are
    independent; only the MMA row a sequence sits in changes)."""
    import mgr_b200 as mgr
    B, T = 128, 24
    monkeypatch.setattr(mgr.models, "SPLIT_MIN_HALF", 64)
    sp, sk, fu = _small_nets(mgr, cuda, Ha=64, Hs=48)
    g = torch.Generator().manual_seed(5)
    xa = torch.randn(B, T, 39, generator=g).to(cuda)
    xs = torch.randn(B, T, 20, generator=g).to(cuda)
    reg = fu.sample_regularisers(B, T, seed=7, step=0, device=cuda)
    monkeypatch.setenv("GR_TOWER_SPLIT", "0")
    whole = fu.merged(xa, xs, reg).clone()
    monkeypatch.setenv("GR_TOWER_SPLIT", "1")
    split = fu.merged(xa, xs, reg).clone()
    torch.cuda.synchronize()
    monkeypatch.setenv("GR_TOWER_SPLIT", "2")
    split2 = fu.merged(xa, xs, reg).clone()
    torch.cuda.synchronize()
    assert (whole - split).abs().max().item() <= 1e-5
    assert (whole - split2).abs().max().item() <= 1e-5
    assert whole.abs().max().item() > 1e-3


def test_pipelined_trainer_equals_serial_steps(cuda):
    """FusionTrainer with the towers one batch ahead must give the same losses and the same weights as the
    plain step-by-step loop (same per-step regulariser streams; the towers are frozen)."""
    import copy
    import mgr_b200 as mgr
    rng = np.random.default_rng(77)
    B, T, C = 4, 30, 22
    _, _, fu1 = _small_nets(mgr, cuda)
    fu2 = copy.deepcopy(fu1)
    batches = []
    for _ in range(4):
        xa = torch.tensor(rng.standard_normal((B, T, 39)).astype(np.float32), device=cuda)
        xs = torch.tensor(rng.standard_normal((B, T, 20)).astype(np.float32), device=cuda)
        labels, ll = random_labels(rng, B, 6, C)
        batches.append((xa, xs, torch.tensor(labels), torch.tensor(np.full((B, 1), T - 2)), torch.tensor(ll)))
    t1 = mgr.FusionTrainer(fu1, mgr.fusion_optimizer(fu1), seed=9, global_batch=B)
    t2 = mgr.FusionTrainer(fu2, mgr.fusion_optimizer(fu2), seed=9, global_batch=B)
    for n, b in enumerate(batches):
        l1 = t1.step(b)                                                        # serial
        nxt = batches[n + 1][:2] if n + 1 < len(batches) else None
        l2 = t2.step(b, next_inputs=nxt)                                       # towers of batch n+1 in flight
        assert torch.allclose(l1, l2, rtol=1e-6, atol=0)
    # (not bit-equal: the weight-gradient GEMMs accumulate split-K tiles with fp32 atomics)
    for p1, p2 in zip(fu1.trainable_parameters(), fu2.trainable_parameters()):
        assert torch.allclose(p1, p2, rtol=0, atol=1e-6)


def test_adam_clip_maxnorm_matches_keras_formula(cuda):
    import mgr_b200 as mgr
    rng = np.random.default_rng(5)
    w = (rng.standard_normal((40, 16)) * 2).astype(np.float32)
    p = torch.tensor(w, device=cuda)
    opt = mgr.KerasAdam([p], lr=1e-2, clipvalue=0.5, decay=1e-3, maxnorm_params=[p], max_norm=3.0)
    ref, m, v = w.astype(np.float64), np.zeros_like(w, dtype=np.float64), np.zeros_like(w, dtype=np.float64)
    for it in range(3):
        g = rng.standard_normal(w.shape).astype(np.float32)
        opt.step([torch.tensor(g, device=cuda)])
        lr = 1e-2 / (1 + 1e-3 * it)
        t = it + 1
        lr_t = lr * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        gc = np.clip(g.astype(np.float64), -0.5, 0.5)
        m = 0.9 * m + 0.1 * gc
        v = 0.999 * v + 0.001 * gc * gc
        ref = ref - lr_t * m / (np.sqrt(v) + 1e-7)
        n = np.sqrt((ref ** 2).sum(0, keepdims=True))
        ref = ref * np.clip(n, 0, 3.0) / (1e-7 + n)
    assert np.abs(p.cpu().numpy() - ref).max() < 1e-5


def test_regularisers_statistics(cuda):
    from mgr_b200 import ops
    m = ops.dropout_mask((8, 64, 1000), 0.4, seed=1, offset=0, device=cuda)
    keep = (m > 0).float().mean().item()
    assert abs(keep - 0.6) < 0.01 and abs(m.max().item() - 1 / 0.6) < 1e-6
    z = ops.gaussian_noise((256, 1000), 0.5, seed=2, offset=0, device=cuda)
    assert abs(z.mean().item()) < 0.01 and abs(z.std().item() - 0.5) < 0.01
    m2 = ops.dropout_mask((8, 64, 1000), 0.4, seed=1, offset=64, device=cuda)
    assert not torch.equal(m, m2)


@pytest.mark.parametrize("R,Fin,C,masked,warp", [(1000, 200, 22, True, False), (777, 1000, 44, False, False),
                                                 (513, 600, 22, True, True), (300, 36, 5, False, False)])
def test_dense_softmax_kernels(cuda, monkeypatch, R, Fin, C, masked, warp):
    """Dense(C)+softmax (speech_lstm_ctc_words.py:86-90): thread-per-row and warp-per-row kernels vs fp64."""
    from mgr_b200 import ops
    if warp:
        monkeypatch.setenv("GR_DENSE_WARP", "1")
    g = torch.Generator().manual_seed(R + C)
    x = torch.randn(R, Fin, generator=g).to(cuda)
    W = (torch.randn(Fin, C, generator=g) * 0.1).to(cuda)
    b = torch.randn(C, generator=g).to(cuda)
    m = ((torch.rand(R, Fin, generator=g) > 0.5).float() * 2).to(cuda) if masked else None
    logits, probs = ops.dense_softmax_fwd(x, W, b, m)
    xd = x.double() * (m.double() if masked else 1.0)
    ref = xd @ W.double() + b.double()
    assert (logits.double() - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
    assert (probs.double() - torch.softmax(ref, -1)).abs().max().item() <= 1e-5


def test_flat_bucket_update_equals_per_tensor_update(cuda):
    """The fused multi-tensor epilogue (pack -> [all-reduce] -> clipvalue + Adam in one launch -> maxnorm) gives the
    per-tensor updates bit for bit, over three steps with lr decay."""
    import mgr_b200 as mgr
    from mgr_b200 import parallel
    g = torch.Generator().manual_seed(3)
    shapes = [(40, 32), (2, 4, 16), (32,), (8, 5), (5,)]
    pa = [torch.nn.Parameter(torch.randn(s, generator=g).to(cuda) * 2) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = mgr.KerasAdam(pa, lr=1e-2, clipvalue=0.5, decay=1e-3, maxnorm_params=[pa[0]], max_norm=3.0)
    ob = mgr.KerasAdam(pb, lr=1e-2, clipvalue=0.5, decay=1e-3, maxnorm_params=[pb[0]], max_norm=3.0)
    bucket = parallel.FlatGradBucket(pb)
    for step in range(3):
        grads = [torch.randn(s, generator=g).to(cuda) for s in shapes]
        oa.step(grads)
        bucket.pack(grads)
        views = bucket.all_reduce()               # single process: no collective, the flat views come back
        assert isinstance(views, parallel.FlatViews) and views.flat is bucket.flat
        ob.step(views)
    for a, b, ma, mb in zip(pa, pb, oa.m, ob.m):
        assert torch.equal(a.data, b.data) and torch.equal(ma, mb)
