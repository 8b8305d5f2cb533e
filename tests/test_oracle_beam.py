"""Pins for the bit-exact C beam oracle (oracle/beam_ref.c) and the deterministic math it shares
with the CUDA kernel (csrc/det_math.h): accuracy vs libm, agreement with the fp64 Python
restatement, exhaustive most-probable-labelling search."""
import math

import numpy as np

from oracle import beam_c, decode_ref
from helpers import peaky_probs


def test_det_math_accuracy():
    rng = np.random.default_rng(0)
    for x in -np.abs(rng.standard_normal(3000) * 25).astype(np.float32):
        ref = math.exp(float(x))
        if ref > 1e-36:
            assert abs(beam_c.det("expf", float(x)) - ref) <= 4e-7 * ref
    assert beam_c.det("expf", -200.0) == 0.0 and beam_c.det("expf", 0.0) == 1.0
    for y in rng.random(3000).astype(np.float32):
        assert abs(beam_c.det("log1pf", float(y)) - math.log1p(float(y))) <= 3e-7 * max(math.log1p(float(y)), 1e-3)
    for z in (10.0 ** rng.uniform(-30, 3, 3000)).astype(np.float32):
        assert abs(beam_c.det("logf", float(z)) - math.log(float(z))) <= 3e-7 * max(abs(math.log(float(z))), 1.0)
    assert beam_c.det("lse", float("-inf"), -3.0) == -3.0
    assert abs(beam_c.det("lse", -1.0, -1.0) - (-1.0 + math.log(2.0))) < 2e-7


def test_c_oracle_matches_python_oracle():
    rng = np.random.default_rng(4001)
    s = peaky_probs(rng, 4, 120, 22)
    for j in range(4):
        for merge in (True, False):
            a = beam_c.beam_search(s[j], 120, 100, 1, merge)
            b = decode_ref.ctc_beam_search(s[j], 120, 100, merge_repeated=merge)
            assert a[0][0] == b[0][0]
            assert abs(float(a[0][1]) - b[0][1]) < 1e-3
    # narrow beams exercise eviction / re-activation
    for W in (1, 2, 5):
        a = beam_c.beam_search(s[0], 60, W, 1, True)
        b = decode_ref.ctc_beam_search(s[0], 60, W, merge_repeated=True)
        assert a[0][0] == b[0][0]


def test_c_oracle_exact_when_wide_and_top_paths():
    rng = np.random.default_rng(12)
    for trial in range(20):
        T, C = int(rng.integers(2, 6)), 3
        q = rng.random((T, C)) ** 2
        q /= q.sum(1, keepdims=True)
        bl, blp, tot = decode_ref.brute_force_best_labelling(q)
        r = beam_c.beam_search(q.astype(np.float32), T, 200, 3, False)
        assert r[0][0] == bl and abs(float(r[0][1]) - blp) < 1e-4
        ranked = sorted(tot.items(), key=lambda kv: -kv[1])
        for k in range(min(3, len(ranked))):
            assert abs(math.exp(float(r[k][1])) - ranked[k][1]) < 1e-5
