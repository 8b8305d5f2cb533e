"""The reference's fusion flow end to end on the GPU path (SURVEY.md 8f ranks 3 + 4; multimodal.py:58-269):
towers rebuilt from JSON + Keras-layout `.h5` weight files, `FusionDataGenerator` batches of an on-disk dataset through
`to_device` into `FusionTrainer.step` (towers one batch ahead), the first step's loss against the fp64 oracle with the
trainer's own regulariser draw, the decoded output through `decode_batch`."""
import numpy as np
import pytest
import torch

from test_data_generator import _write_dataset

pytestmark = pytest.mark.gpu


def _t64(ws):
    return [torch.tensor(w, dtype=torch.float64) for w in ws]


def test_generator_h5_towers_trainer_and_decode(cuda, tmp_path):
    import mgr_b200 as mgr
    from mgr_b200 import keras_io
    from mgr_b200.data_generator import FusionDataGenerator, to_device
    from oracle import lstm_ref
    rng = np.random.default_rng(5)
    root = tmp_path / "data"
    root.mkdir()
    _write_dataset(str(root), "train", list(range(200, 216)), rng, blank_ids=(203,), no_skeletal=(207,))
    # "previously trained" uni-modal networks, saved the way the reference saves them (data_generator.py:277-281)
    sp0 = mgr.UnimodalNet(39, 16, 44, 0.5, (0.4, 0.5, 0.5), seed=21)
    sk0 = mgr.UnimodalNet(20, 12, 22, 0.5, (0.6, 0.6, 0.6), seed=22)
    files = {}
    for name, net in (("sp", sp0), ("sk", sk0)):
        files[name] = (str(tmp_path / ("%s_ctc_lstm_model.json" % name)), str(tmp_path / ("%s_ctc_lstm_weights_best.h5" % name)))
        open(files[name][0], "w").write(keras_io.to_json(net))
        keras_io.save_weights(net, files[name][1])
    # multimodal.py:68-85: model_from_json + load_weights, then the fusion model on top of the (frozen) towers
    towers = []
    for name in ("sp", "sk"):
        net = keras_io.model_from_json(open(files[name][0]).read())
        keras_io.load_weights(net, files[name][1])
        towers.append(net)
    fu = mgr.FusionNet(towers[0], towers[1], nb_classes=22, units=8, seed=23).to(cuda)
    for a, b in zip(fu.speech.blstm_2.get_weights(), sp0.blstm_2.get_weights()):
        assert np.array_equal(a, b)
    B, T = 4, 26
    gen = FusionDataGenerator(B, 20, 39, T, 22, "train", data_root=str(root))
    it = gen.next_train()
    host = [next(it)[0] for _ in range(3)]
    dev = [to_device(h, cuda) for h in host]
    opt = mgr.fusion_optimizer(fu)
    trainer = mgr.FusionTrainer(fu, opt, seed=31, global_batch=B)
    w_before = _t64(fu.blstm_3.get_weights()), _t64(fu.dense.get_weights())
    losses = []
    for s in range(3):
        nxt = (dev[s + 1][0], dev[s + 1][1]) if s + 1 < 3 else None
        losses.append(trainer.step(dev[s], next_inputs=nxt).cpu().numpy())
    trainer.close()
    assert all(np.isfinite(l).all() for l in losses)
    # step 0 against the oracle, with the regularisers the trainer drew for step 0 (Philox: same seed + step -> same draw)
    reg = fu.sample_regularisers(B, T, seed=31, step=0, device=cuda)
    m3 = reg["m3"].cpu().double()
    masks = {"fu_f": m3[:4], "fu_b": m3[4:]}
    for pre, t in (("sp_", "sp"), ("sk_", "sk")):
        for key, name in (("m1", "l1"), ("m2", "l2")):
            m = reg[t][key].cpu().double()
            masks[pre + name + "f"], masks[pre + name + "b"] = m[:4], m[4:]
    xa, xs = torch.tensor(host[0]["the_input_audio"], dtype=torch.float64), torch.tensor(host[0]["the_input_skeletal"], dtype=torch.float64)
    p, _, _ = lstm_ref.fusion_forward(xa, xs, _t64(sp0.blstm_1.get_weights()), _t64(sp0.blstm_2.get_weights()),
                                      _t64(sk0.blstm_1.get_weights()), _t64(sk0.blstm_2.get_weights()), w_before[0], w_before[1],
                                      noise_a=reg["sp"]["noise"].cpu().double(), masks=masks, drop_mask=reg["drop"].cpu().double())
    ref = lstm_ref.torch_ctc_lambda(p, host[0]["the_labels"], host[0]["input_length"], host[0]["label_length"]).numpy()[:, 0]
    assert np.abs(losses[0] - ref).max() <= 1e-4 * np.abs(ref).max()
    # the trained model decodes the validation batch into the reference's label strings + MLF
    val = next(gen.next_val())[0]
    xa_v, xs_v = to_device(val, cuda)[:2]
    with torch.no_grad():
        y_pred, _ = fu(xa_v, xs_v)
    out = mgr.decode_batch(y_pred, list(range(1, y_pred.shape[0] + 1)), mlf_path=str(tmp_path / "final_ctc_recout.mlf"))
    assert len(out) == y_pred.shape[0] and all(isinstance(s, str) for seq in out for s in seq)
    assert (tmp_path / "final_ctc_recout.mlf").read_text().startswith("#!MLF!#\n")
