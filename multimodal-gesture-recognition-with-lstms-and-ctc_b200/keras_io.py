"""Keras-layout persistence for the three networks: `to_json` / `model_from_json` and
`save_weights` / `load_weights`, mirroring what the reference does around training
(`model.to_json()` + `model.save_weights(...)` in DataGenerator.on_epoch_end,
/root/reference/audio_network/data_generator.py:277-281; `model_from_json` + `load_weights` of the two
towers, /root/reference/multimodal_fusion/multimodal.py:68-85) and the tower re-use
`speech_model.layers[2](x)`, `.layers[3](...)` (`multimodal.py:109-118`).

Weights travel in KERAS ORDER -- per layer in `model.layers` order, per layer in `layer.weights` order,
i.e. for `Bidirectional(LSTM)`: forward kernel (F,4H), forward recurrent kernel (H,4H), forward bias (4H),
then the backward three; gate order i,f,c,o; for `Dense`: kernel, bias -- under the names Keras 2.1.4
gives them (`bidirectional_1/forward_blstm_1/kernel:0`, ...), which is also how `load_weights` matches:
by position, names are informative (Keras' topological loading).

Containers: `.h5` / `.hdf5` -- the reference's own format, through the pure-Python reader / writer of
`keras_h5.py` (classic HDF5 as h5py writes it for Keras 2.1.4: `layer_names` / `weight_names` attributes, one
dataset per weight) -- and `.npz` (NumPy), which `scripts/h5_to_npz.py` also produces on a machine with h5py.
"""
import json
from collections import OrderedDict

import numpy as np

from .models import UnimodalNet, FusionNet

_BLSTM_PARTS = ("kernel:0", "recurrent_kernel:0", "bias:0")


def _blstm_entries(wrapper_name, lstm_name, layer):
    w = layer.get_weights()
    out = []
    for d, direction in enumerate(("forward", "backward")):
        for k, part in enumerate(_BLSTM_PARTS):
            out.append(("%s/%s_%s/%s" % (wrapper_name, direction, lstm_name, part), w[3 * d + k]))
    return out


def weight_table(model, prefix=""):
    """OrderedDict layer name -> [(weight name, array)] in Keras saving order."""
    t = OrderedDict()
    if isinstance(model, UnimodalNet):
        t[prefix + "bidirectional_1"] = _blstm_entries(prefix + "bidirectional_1", "blstm_1", model.blstm_1)
        t[prefix + "bidirectional_2"] = _blstm_entries(prefix + "bidirectional_2", "blstm_2", model.blstm_2)
        d = model.dense.get_weights()
        t[prefix + "dense_1"] = [(prefix + "dense_1/kernel:0", d[0]), (prefix + "dense_1/bias:0", d[1])]
        return t
    if isinstance(model, FusionNet):
        # Keras 2.1.4 `Container.layers` is sorted by DEPTH (distance from the output), ties in the order the graph
        # walk from the outputs first met the layer.  In multimodal.py:109-118 both towers hang off the Merge at the
        # same depths, so the saved fusion file (and topological `load_weights`) interleaves them:
        #   speech_blstm_1 (depth 8), skeletal_blstm_1 (8), speech_blstm_2 (7), skeletal_blstm_2 (7),
        #   the new Bidirectional (4), dense_1 (2)   -- layer names as renamed at multimodal.py:121-128;
        # the uni-modal heads are not part of the fusion graph.
        for lname, lstm, layer in (("speech_blstm_1", "blstm_1", model.speech.blstm_1),
                                   ("skeletal_blstm_1", "blstm_1", model.skeletal.blstm_1),
                                   ("speech_blstm_2", "blstm_2", model.speech.blstm_2),
                                   ("skeletal_blstm_2", "blstm_2", model.skeletal.blstm_2)):
            t[lname] = _blstm_entries(lname, lstm, layer)
        t["bidirectional_3"] = _blstm_entries("bidirectional_3", "blstm_2", model.blstm_3)
        d = model.dense.get_weights()
        t["dense_1"] = [("dense_1/kernel:0", d[0]), ("dense_1/bias:0", d[1])]
        return t
    raise TypeError("weight_table: unsupported model %r" % type(model))


def _assign(model, arrays):
    """Inverse of weight_table's flattening: consume `arrays` in Keras order."""
    it = iter(arrays)

    def take(n):
        return [next(it) for _ in range(n)]
    if isinstance(model, UnimodalNet):
        model.blstm_1.set_weights(take(6))
        model.blstm_2.set_weights(take(6))
        model.dense.set_weights(take(2))
    else:
        # depth order of the fusion graph (see weight_table): s1, k1, s2, k2
        model.speech.blstm_1.set_weights(take(6))
        model.skeletal.blstm_1.set_weights(take(6))
        model.speech.blstm_2.set_weights(take(6))
        model.skeletal.blstm_2.set_weights(take(6))
        model.blstm_3.set_weights(take(6))
        model.dense.set_weights(take(2))
    rest = list(it)
    if rest:
        raise ValueError("load_weights: %d arrays left over" % len(rest))


def _is_h5(path):
    return str(path).lower().endswith((".h5", ".hdf5"))


def save_weights(model, path):
    """`model.save_weights(path)`.  `.h5`: the Keras 2.1.4 HDF5 tree (keras_h5.write_keras_weights); otherwise `.npz`,
    one array per Keras weight, keys 'NNN|<keras weight name>' (NNN keeps the order)."""
    if _is_h5(path):
        from . import keras_h5
        table = weight_table(model)
        keras_h5.write_keras_weights(path, [(lname, entries) for lname, entries in table.items()])
        return [n for entries in table.values() for n, _ in entries]
    flat = [(n, a) for entries in weight_table(model).values() for n, a in entries]
    np.savez(path, **{"%03d|%s" % (i, n): a for i, (n, a) in enumerate(flat)})
    return [n for n, _ in flat]


def load_weights(model, path):
    """`model.load_weights(path)`: topological (positional) matching with shape checks, like Keras.  `.h5` files are read
    in `layer_names` / `weight_names` order (what load_weights_from_hdf5_group iterates)."""
    if _is_h5(path):
        from . import keras_h5
        pairs = [(n, a) for _, ws in keras_h5.read_keras_weights(path) for n, a in ws]
        keys, arrays = [n for n, _ in pairs], [np.asarray(a, dtype=np.float32) for _, a in pairs]
    else:
        with np.load(path) as z:
            keys = sorted(z.files, key=lambda k: int(k.split("|", 1)[0]))
            arrays = [z[k] for k in keys]
    want = [a for entries in weight_table(model).values() for _, a in entries]
    if len(arrays) != len(want):
        raise ValueError("load_weights: file holds %d arrays, the model needs %d" % (len(arrays), len(want)))
    for k, a, w in zip(keys, arrays, want):
        if a.shape != w.shape:
            raise ValueError("load_weights: %s has shape %s, the model expects %s" % (k, a.shape, w.shape))
    _assign(model, arrays)
    return keys


def _unimodal_config(net):
    return {"class_name": "UnimodalNet",
            "config": {"numfeats": net.numfeats, "units": net.units, "nb_classes": net.nb_classes,
                       "noise_std": net.noise_std, "dropouts": [net.p1, net.p2, net.pd]},
            # what the reference's model.to_json() would describe (speech_lstm_ctc_words.py:46-90)
            "keras_layers": ["InputLayer", "GaussianNoise", "Bidirectional(LSTM blstm_1)", "Bidirectional(LSTM blstm_2)",
                             "Add", "Dropout", "Dense dense_1", "Activation softmax"]}


def to_json(model):
    """`model.to_json()`: the topology (not the weights) as a JSON string `model_from_json` rebuilds."""
    if isinstance(model, UnimodalNet):
        return json.dumps(_unimodal_config(model))
    if isinstance(model, FusionNet):
        return json.dumps({"class_name": "FusionNet",
                           "config": {"nb_classes": model.nb_classes, "units": model.units,
                                      "speech": _unimodal_config(model.speech),
                                      "skeletal": _unimodal_config(model.skeletal)}})
    raise TypeError("to_json: unsupported model %r" % type(model))


def model_from_json(s):
    """`keras.models.model_from_json`: rebuild the network (fresh initial weights; call load_weights next)."""
    d = json.loads(s)

    def uni(c):
        c = c["config"]
        return UnimodalNet(c["numfeats"], c["units"], c["nb_classes"], c["noise_std"], tuple(c["dropouts"]))
    if d["class_name"] == "UnimodalNet":
        return uni(d)
    if d["class_name"] == "FusionNet":
        c = d["config"]
        return FusionNet(uni(c["speech"]), uni(c["skeletal"]), nb_classes=c["nb_classes"], units=c["units"])
    raise ValueError("model_from_json: unknown class %r" % d["class_name"])
