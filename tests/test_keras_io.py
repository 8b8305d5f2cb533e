"""Keras-order weight / topology persistence, CPU only: nothing here launches a kernel."""
import numpy as np
import pytest


def _nets():
    import mgr_b200 as mgr
    sp = mgr.UnimodalNet(39, 8, 44, 0.5, (0.4, 0.5, 0.5), seed=1)
    sk = mgr.UnimodalNet(20, 6, 22, 0.0, (0.6, 0.6, 0.6), seed=2)
    return sp, sk, mgr.FusionNet(sp, sk, nb_classes=22, units=4, seed=3)


def test_weight_names_and_order_follow_keras():
    from mgr_b200 import keras_io
    sp, _, fu = _nets()
    t = keras_io.weight_table(sp)
    assert list(t) == ["bidirectional_1", "bidirectional_2", "dense_1"]
    names = [n for n, _ in t["bidirectional_1"]]
    assert names == ["bidirectional_1/forward_blstm_1/kernel:0", "bidirectional_1/forward_blstm_1/recurrent_kernel:0",
                     "bidirectional_1/forward_blstm_1/bias:0", "bidirectional_1/backward_blstm_1/kernel:0",
                     "bidirectional_1/backward_blstm_1/recurrent_kernel:0", "bidirectional_1/backward_blstm_1/bias:0"]
    shapes = [a.shape for _, a in t["bidirectional_1"]]
    assert shapes == [(39, 32), (8, 32), (32,), (39, 32), (8, 32), (32,)]       # (F,4H), (H,4H), (4H) per direction
    # unit_forget_bias: the f-gate quarter of each bias is 1 (gate order i,f,c,o)
    b = t["bidirectional_1"][2][1]
    assert np.all(b[8:16] == 1) and np.all(b[:8] == 0) and np.all(b[16:] == 0)
    tf = keras_io.weight_table(fu)
    # Keras depth order of the fusion graph: the two towers interleave (multimodal.py:109-128)
    assert list(tf) == ["speech_blstm_1", "skeletal_blstm_1", "speech_blstm_2", "skeletal_blstm_2",
                        "bidirectional_3", "dense_1"]
    assert tf["bidirectional_3"][0][1].shape == (2 * 8 + 2 * 6, 16)


def test_save_load_roundtrip_and_tower_reuse(tmp_path):
    import mgr_b200 as mgr
    from mgr_b200 import keras_io
    sp, sk, fu = _nets()
    p_sp, p_sk, p_fu = (str(tmp_path / n) for n in ("sp.npz", "sk.npz", "fu.npz"))
    keras_io.save_weights(sp, p_sp)
    keras_io.save_weights(sk, p_sk)
    keras_io.save_weights(fu, p_fu)
    # multimodal.py:68-85: rebuild the towers from JSON + weights, then the fusion model on top of them
    sp2 = keras_io.model_from_json(keras_io.to_json(sp))
    sk2 = keras_io.model_from_json(keras_io.to_json(sk))
    keras_io.load_weights(sp2, p_sp)
    keras_io.load_weights(sk2, p_sk)
    for a, b in zip(sp.blstm_2.get_weights() + sp.dense.get_weights(), sp2.blstm_2.get_weights() + sp2.dense.get_weights()):
        assert np.array_equal(a, b)
    fu2 = mgr.FusionNet(sp2, sk2, nb_classes=22, units=4, seed=99)
    keras_io.load_weights(fu2, p_fu)
    for (n1, a), (n2, b) in zip([e for v in keras_io.weight_table(fu).values() for e in v],
                                [e for v in keras_io.weight_table(fu2).values() for e in v]):
        assert n1 == n2 and np.array_equal(a, b)
    # the towers inside the fusion model stay frozen, the new layers trainable (multimodal.py:135-148)
    assert not fu2.speech.blstm_1.kernel.requires_grad and fu2.blstm_3.kernel.requires_grad
    fu3 = keras_io.model_from_json(keras_io.to_json(fu))
    assert fu3.units == 4 and fu3.speech.units == 8 and fu3.skeletal.numfeats == 20


def test_load_rejects_wrong_topology(tmp_path):
    from mgr_b200 import keras_io
    sp, sk, fu = _nets()
    p = str(tmp_path / "sp.npz")
    keras_io.save_weights(sp, p)
    with pytest.raises(ValueError):
        keras_io.load_weights(sk, p)         # same number of arrays, different shapes
    with pytest.raises(ValueError):
        keras_io.load_weights(fu, p)         # different number of arrays


def test_fusion_file_laid_out_as_keras_writes_it(tmp_path):
    """A fusion checkpoint as Keras 2.1.4 saves it: layers in DEPTH order -- speech_blstm_1, skeletal_blstm_1,
    speech_blstm_2, skeletal_blstm_2, bidirectional_3, dense_1 -- each Bidirectional as forward (kernel, recurrent,
    bias) then backward.  Built by hand here (not through save_weights) and loaded positionally."""
    from mgr_b200 import keras_io
    sp, sk, fu = _nets()
    rng = np.random.default_rng(5)
    Hs, Hk, Hf = 8, 6, 4

    def blstm(F, H):
        return [rng.standard_normal(s).astype(np.float32) for s in ((F, 4 * H), (H, 4 * H), (4 * H,)) * 2]
    layers = {"speech_blstm_1": blstm(39, Hs), "skeletal_blstm_1": blstm(20, Hk), "speech_blstm_2": blstm(2 * Hs, Hs),
              "skeletal_blstm_2": blstm(2 * Hk, Hk), "bidirectional_3": blstm(2 * Hs + 2 * Hk, Hf),
              "dense_1": [rng.standard_normal((2 * Hf, 22)).astype(np.float32), rng.standard_normal(22).astype(np.float32)]}
    order = ["speech_blstm_1", "skeletal_blstm_1", "speech_blstm_2", "skeletal_blstm_2", "bidirectional_3", "dense_1"]
    flat = [(n, a) for n in order for a in layers[n]]
    path = str(tmp_path / "fusion_as_keras.npz")
    np.savez(path, **{"%03d|%s/w%d" % (i, n, i): a for i, (n, a) in enumerate(flat)})
    keras_io.load_weights(fu, path)
    for got, want in ((fu.speech.blstm_1, "speech_blstm_1"), (fu.skeletal.blstm_1, "skeletal_blstm_1"),
                      (fu.speech.blstm_2, "speech_blstm_2"), (fu.skeletal.blstm_2, "skeletal_blstm_2"),
                      (fu.blstm_3, "bidirectional_3"), (fu.dense, "dense_1")):
        for a, b in zip(got.get_weights(), layers[want]):
            assert np.array_equal(a, b)
    # the tower-major order (speech 1, 2, skeletal 1, 2) is NOT what Keras writes: first mismatch (16,32) vs (20,24)
    bad = [(n, a) for n in ["speech_blstm_1", "speech_blstm_2", "skeletal_blstm_1", "skeletal_blstm_2", "bidirectional_3",
                            "dense_1"] for a in layers[n]]
    path2 = str(tmp_path / "tower_major.npz")
    np.savez(path2, **{"%03d|%s/w%d" % (i, n, i): a for i, (n, a) in enumerate(bad)})
    with pytest.raises(ValueError):
        keras_io.load_weights(fu, path2)
