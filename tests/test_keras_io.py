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
    assert list(tf) == ["speech/bidirectional_1", "speech/bidirectional_2", "skeletal/bidirectional_1",
                        "skeletal/bidirectional_2", "bidirectional_3", "dense_1"]
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
