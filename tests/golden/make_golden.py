"""Generates the committed golden fixtures under tests/golden/ from the ORACLE (the reference
itself cannot run here: Python 2 + Keras 2.1.4/TF 1.12.1; it ships no fixtures).
Run from the repo root:  python tests/golden/make_golden.py
The fixtures pin (a) the oracle against regressions (CPU suite) and (b) the CUDA path (GPU suite)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ctc_ref, decode_ref, lstm_ref, beam_c  # noqa: E402
from helpers import random_probs, random_labels, peaky_probs  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def ctc_case():
    rng = np.random.default_rng(3001)
    B, T, C, Lmax = 6, 42, 22, 12
    p, a = random_probs(rng, B, T, C)
    il = rng.integers(20, T - 1, size=(B, 1))
    labels, ll = random_labels(rng, B, Lmax, C, T_avail=il[:, 0])
    labels[1, :ll[1, 0]] = C - 1  # the reference's "blank example" (data_generator.py:207-213)
    loss, g_p = ctc_ref.ctc_lambda_func((p, labels, il, ll), want_grad=True)
    loss_a, g_a = ctc_ref.softmax_ctc_grad_logits(a, labels, il, ll)
    np.savez_compressed(os.path.join(OUT, "ctc_case.npz"), probs=p, logits=a, labels=labels, input_length=il,
                        label_length=ll, loss=loss, grad_probs=g_p, loss_from_logits=loss_a, grad_logits=g_a)


def decode_case():
    rng = np.random.default_rng(4001)
    N, T, C = 5, 90, 22
    s = peaky_probs(rng, N, T, C, sharp=2.5)
    out = {"probs": s}
    for thr in (0.5, 0.75, 0.97):
        ids = [decode_ref.decode_ids_literal(s[j], thr) for j in range(N)]
        m = -np.ones((N, T), dtype=np.int32)
        for j, v in enumerate(ids):
            m[j, :len(v)] = v
        out["ids_thr_%d" % int(thr * 100)] = m
    g = -np.ones((N, T), dtype=np.int32)
    bm = -np.ones((N, T), dtype=np.int32)
    blp = np.zeros(N, dtype=np.float32)
    for j in range(N):
        v, _ = decode_ref.ctc_greedy(s[j], T)
        g[j, :len(v)] = v
        r = beam_c.beam_search(s[j], T, 100, 1, True)
        bm[j, :len(r[0][0])] = r[0][0]
        blp[j] = r[0][1]
    out["greedy"] = g
    out["beam100"] = bm
    out["beam100_logp"] = blp
    np.savez_compressed(os.path.join(OUT, "decode_case.npz"), **out)


def blstm_case():
    rng = np.random.default_rng(47)
    B, T, F, H = 3, 14, 39, 20
    w6 = lstm_ref.init_blstm_weights(rng, F, H)
    x = rng.standard_normal((B, T, F)).astype(np.float32)
    masks = ((rng.random((8, B, F)) > 0.4) / 0.6).astype(np.float32)
    dy = rng.standard_normal((B, T, 2 * H)).astype(np.float32)
    t64 = [torch.tensor(w, dtype=torch.float64, requires_grad=True) for w in w6]
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    m = torch.tensor(masks, dtype=torch.float64)
    y = lstm_ref.bidirectional_lstm(xt, t64, m[:4], m[4:])
    (y * torch.tensor(dy, dtype=torch.float64)).sum().backward()
    np.savez_compressed(os.path.join(OUT, "blstm_case.npz"), x=x, masks=masks, dy=dy, y=y.detach().numpy(),
                        dx=xt.grad.numpy(), **{"w%d" % i: w6[i] for i in range(6)},
                        **{"dw%d" % i: t64[i].grad.numpy() for i in range(6)})


if __name__ == "__main__":
    ctc_case()
    decode_case()
    blstm_case()
    print("golden fixtures written to", OUT)
