"""Shared synthetic-input generators for the parity tests (seeded; SURVEY.md 8d conventions)."""
import numpy as np


def random_probs(rng, B, T, C, scale=2.0):
    a = rng.standard_normal((B, T, C)) * scale
    e = np.exp(a - a.max(axis=2, keepdims=True))
    return (e / e.sum(axis=2, keepdims=True)).astype(np.float32), a.astype(np.float32)


def random_labels(rng, B, Lmax, C, lmin=1, lmax=None, T_avail=None):
    lmax = Lmax if lmax is None else lmax
    labels = -np.ones((B, Lmax), dtype=np.float32)
    ll = np.zeros((B, 1), dtype=np.int64)
    for b in range(B):
        L = int(rng.integers(lmin, lmax + 1))
        if T_avail is not None:
            L = max(1, min(L, int(T_avail[b]) // 2))
        labels[b, :L] = rng.integers(0, C - 1, size=L)
        ll[b, 0] = L
    return labels, ll


def peaky_probs(rng, N, T, C, sharp=4.0):
    """Piecewise-constant class tracks + noise (config 5 of BASELINE.json)."""
    out = np.zeros((N, T, C), dtype=np.float32)
    for n in range(N):
        t = 0
        track = np.zeros(T, dtype=np.int64)
        while t < T:
            seg = int(rng.integers(5, 61))
            c = C - 1 if rng.random() < 0.5 else int(rng.integers(0, C - 1))
            track[t:t + seg] = c
            t += seg
        logits = rng.standard_normal((T, C))
        logits[np.arange(T), track] += sharp
        e = np.exp(logits - logits.max(axis=1, keepdims=True))
        out[n] = (e / e.sum(axis=1, keepdims=True)).astype(np.float32)
    return out
