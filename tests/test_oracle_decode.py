"""Pins for oracle/decode_ref.py: literal Python-2 loop vs closed form (property test), greedy
rules, and beam == exhaustive most-probable labelling when the beam is wide enough."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import decode_ref


@settings(max_examples=150, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(3, 90), st.integers(2, 9), st.sampled_from([0.5, 0.75, 0.97, 0.0]))
def test_literal_equals_closed_form(seed, T, C, thr):
    rng = np.random.default_rng(seed)
    s = rng.random((T, C)).astype(np.float32) ** 3
    s /= s.sum(1, keepdims=True)
    assert decode_ref.decode_ids_literal(s, thr) == decode_ref.decode_ids_closed_form(s, thr)


def test_filter_is_count_based_not_positional():
    # class 1 at frames 2,3(low),4 ; the Python-2 loop removes the FIRST occurrence, not frame 3
    s = np.full((7, 3), 0.01, dtype=np.float32)
    best = [0, 0, 1, 1, 2, 1, 2]
    conf = [.9, .9, .9, .4, .9, .9, .9]
    for t, (b, c) in enumerate(zip(best, conf)):
        s[t, b] = c
    ids = decode_ref.decode_ids_literal(s, 0.5)
    assert ids == [1, 2, 1, 2]
    assert decode_ref.decode_batch(s[None], [1], mlf_path=None) == [["VA", "VQ", "VA", "VQ"]]


def test_mlf_format(tmp_path):
    rng = np.random.default_rng(0)
    s = rng.random((3, 12, 22)).astype(np.float32)
    s /= s.sum(2, keepdims=True)
    path = tmp_path / "out.mlf"
    ret = decode_ref.decode_batch(s, [7, 228, 12], threshold=0.0, mlf_path=str(path))
    txt = path.read_text().splitlines()
    assert txt[0] == "#!MLF!#" and txt[1] == '"*/Sample00007.rec"'
    assert '"*/Sample00228.rec"' not in txt and len(ret) == 3
    assert txt.count(".") == 2


def test_greedy_rules():
    C = 4
    seq = [0, 0, 3, 0, 1, 1, 3, 3, 2]
    p = np.full((len(seq), C), 0.05, dtype=np.float32)
    for t, c in enumerate(seq):
        p[t, c] = 0.85
    out, score = decode_ref.ctc_greedy(p, len(seq))
    assert out == [0, 0, 1, 2]
    assert abs(score + len(seq) * np.log(np.float32(0.85) + 1e-8)) < 1e-4


def test_beam_is_exact_when_wide():
    rng = np.random.default_rng(11)
    ok = 0
    for trial in range(25):
        T, C = int(rng.integers(2, 6)), 3
        s = rng.random((T, C)) ** 2
        s /= s.sum(1, keepdims=True)
        bl, blp, _ = decode_ref.brute_force_best_labelling(s)
        r = decode_ref.ctc_beam_search(s, T, beam_width=200, merge_repeated=False)
        ok += int(r[0][0] == bl and abs(r[0][1] - blp) < 1e-9)
    assert ok == 25
