"""Pins for oracle/ctc_ref.py (the reference ships no fixtures -- SURVEY.md 8c): exhaustive path
enumeration, torch-CPU F.ctc_loss as an independent implementation, fp64 finite differences,
hand-derived known answers, TF error rules."""
import numpy as np
import pytest
import torch

from oracle import ctc_ref, lstm_ref
from helpers import random_probs, random_labels


def test_brute_force_enumeration():
    rng = np.random.default_rng(0)
    for trial in range(12):
        T, C = int(rng.integers(1, 7)), int(rng.integers(2, 5))
        y = rng.random((T, C)) + 0.05
        y /= y.sum(1, keepdims=True)
        L = int(rng.integers(0, min(3, T) + 1))
        lab = list(rng.integers(0, C - 1, size=L)) if C > 1 else []
        loss, _, _, _ = ctc_ref.ctc_loss_grad_single(np.log(y), lab, T)
        bf = ctc_ref.brute_force_neg_log_prob(y, lab)
        if np.isinf(bf):
            assert np.isinf(loss)
        else:
            assert abs(loss - bf) < 1e-12


def test_contract_matches_torch_ctc_loss():
    rng = np.random.default_rng(1)
    B, T, C, Lm = 6, 20, 7, 5
    p, _ = random_probs(rng, B, T, C)
    il = rng.integers(10, T - 1, size=(B, 1))
    labels, ll = random_labels(rng, B, Lm, C, T_avail=il[:, 0])
    loss, g = ctc_ref.ctc_lambda_func((p, labels, il, ll), want_grad=True)
    pt = torch.tensor(p, dtype=torch.float64, requires_grad=True)
    lt = lstm_ref.torch_ctc_lambda(pt, labels, il, ll)
    lt.sum().backward()
    assert np.abs(loss - lt.detach().numpy()).max() < 1e-10
    assert np.abs(g - pt.grad.numpy()).max() < 1e-9
    assert np.all(g[:, :2] == 0)


def test_finite_difference_gradient_wrt_probs():
    rng = np.random.default_rng(2)
    B, T, C = 2, 7, 4
    p, _ = random_probs(rng, B, T, C, scale=1.0)
    p = p.astype(np.float64)
    labels = np.array([[0, 1], [2, -1]], dtype=np.float32)
    ll = np.array([[2], [1]])
    il = np.array([[5], [4]])
    loss, g = ctc_ref.ctc_lambda_func((p, labels, il, ll), want_grad=True)
    h = 1e-6
    for (b, t, c) in [(0, 2, 0), (0, 4, 3), (1, 3, 2), (1, 5, 1), (0, 0, 1)]:
        pp, pm = p.copy(), p.copy()
        pp[b, t, c] += h
        pm[b, t, c] -= h
        fd = (ctc_ref.ctc_lambda_func((pp, labels, il, ll))[b, 0] - ctc_ref.ctc_lambda_func((pm, labels, il, ll))[b, 0]) / (2 * h)
        assert abs(fd - g[b, t, c]) < 1e-6 * max(1.0, abs(fd))


def test_logit_gradient_chain():
    rng = np.random.default_rng(3)
    B, T, C = 3, 9, 5
    _, a = random_probs(rng, B, T, C)
    labels, ll = random_labels(rng, B, 3, C)
    il = np.full((B, 1), T - 2)
    loss, ga = ctc_ref.softmax_ctc_grad_logits(a, labels, il, ll)
    at = torch.tensor(a, dtype=torch.float64, requires_grad=True)
    lt = lstm_ref.torch_ctc_lambda(torch.softmax(at, -1), labels, il, ll)
    lt.mean().backward()
    assert np.abs(ga - at.grad.numpy()).max() < 1e-10


def test_known_answers():
    eps = 1e-8
    # all-blank target: label [C-1] -> empty sequence (data_generator.py:207-213) -> -sum log q(blank)
    T, C = 6, 4
    rng = np.random.default_rng(4)
    p, _ = random_probs(rng, 1, T + 2, C)
    labels = np.array([[C - 1]], dtype=np.float32)
    loss = ctc_ref.ctc_lambda_func((p, labels, np.array([[T]]), np.array([[1]])))
    q = (p[0, 2:] + eps) / (p[0, 2:] + eps).sum(1, keepdims=True)
    assert abs(loss[0, 0] + np.log(q[:, C - 1]).sum()) < 1e-5
    # uniform p, L=1, T frames: #alignments of a single label = T(T+1)/2
    pu = np.full((1, T + 2, C), 1.0 / C)
    loss = ctc_ref.ctc_lambda_func((pu, np.array([[1.0]]), np.array([[T]]), np.array([[1]])))
    assert abs(loss[0, 0] + np.log((T * (T + 1) / 2) / C ** T)) < 1e-6
    # T=1, L=1
    loss = ctc_ref.ctc_lambda_func((pu[:, :3], np.array([[0.0]]), np.array([[1]]), np.array([[1]])))
    assert abs(loss[0, 0] + np.log(1.0 / C)) < 1e-6


def test_tf_error_rules():
    p = np.full((1, 8, 4), 0.25)
    with pytest.raises(ctc_ref.CTCInvalidArgument):
        ctc_ref.ctc_lambda_func((p, np.array([[0.0]]), np.array([[6]]), np.array([[0]])))        # zero labels
    with pytest.raises(ctc_ref.CTCInvalidArgument):
        ctc_ref.ctc_lambda_func((p, np.array([[3.0, 1.0]]), np.array([[6]]), np.array([[2]])))    # non-null after null
    with pytest.raises(ctc_ref.CTCInvalidArgument):
        ctc_ref.ctc_lambda_func((p, np.zeros((1, 7)), np.array([[6]]), np.array([[7]])))          # not enough time
    # repeated labels need an extra frame: no valid path -> +inf, grad = softmax
    p2 = np.full((1, 4, 4), 0.25)
    loss, g = ctc_ref.ctc_batch_cost(np.array([[1.0, 1.0]]), p2[:, :2], np.array([[2]]), np.array([[2]]))
    assert np.isinf(loss[0, 0])


def test_fp32_mode_close_to_fp64():
    rng = np.random.default_rng(5)
    p, _ = random_probs(rng, 2, 60, 22)
    labels, ll = random_labels(rng, 2, 10, 22)
    il = np.full((2, 1), 58)
    l64 = ctc_ref.ctc_lambda_func((p, labels, il, ll))
    l32 = ctc_ref.ctc_lambda_func((p, labels, il, ll), dtype=np.float32)
    assert np.abs(l64 - l32).max() / np.abs(l64).max() < 1e-5
