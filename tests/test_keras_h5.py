"""Pure-Python HDF5 reader / writer for Keras 2.1.4 weight files (multimodal.py:68-85), CPU only.
The sandbox has no h5py, so the reader is checked against (a) the writer's files, (b) hand-assembled byte sequences of the
HDF5 File Format Specification's structures that the writer does NOT emit (attribute message versions 2/3, dataspace
version 2, compact layout, object-header continuation), (c) its refusal of what it does not implement."""
import struct

import numpy as np
import pytest


def _nets():
    import mgr_b200 as mgr
    sp = mgr.UnimodalNet(39, 8, 44, 0.5, (0.4, 0.5, 0.5), seed=1)
    sk = mgr.UnimodalNet(20, 6, 22, 0.0, (0.6, 0.6, 0.6), seed=2)
    return sp, sk, mgr.FusionNet(sp, sk, nb_classes=22, units=4, seed=3)


def test_weight_file_tree_is_what_keras_writes(tmp_path):
    from mgr_b200 import keras_io, keras_h5
    sp, _, _ = _nets()
    path = str(tmp_path / "sp_ctc_lstm_weights_best.h5")
    keras_io.save_weights(sp, path)
    f = keras_h5.H5Reader(path)
    a = f.attrs(f.root)
    assert [x.decode() for x in a["layer_names"]] == ["bidirectional_1", "bidirectional_2", "dense_1"]
    assert a["backend"].tobytes().rstrip(b"\0") == b"tensorflow" and a["keras_version"].tobytes().rstrip(b"\0") == b"2.1.4"
    g = f.resolve(f.root, "bidirectional_1")
    wn = [x.decode() for x in f.attrs(g)["weight_names"]]
    assert wn[0] == "bidirectional_1/forward_blstm_1/kernel:0" and len(wn) == 6
    # the '/' in a weight name makes nested groups, exactly as h5py's create_dataset does
    assert sorted(f.children(g)) == ["bidirectional_1"]
    inner = f.children(f.resolve(g, "bidirectional_1"))
    assert sorted(inner) == ["backward_blstm_1", "forward_blstm_1"]
    k = f.dataset(f.resolve(g, "bidirectional_1/forward_blstm_1/kernel:0"))
    assert k.dtype == np.float32 and k.shape == (39, 32)
    assert np.array_equal(k, sp.blstm_1.get_weights()[0])


def test_h5_roundtrip_and_tower_reuse(tmp_path):
    """multimodal.py:68-85 with the reference's own container: towers from JSON + .h5, fusion model on top, fusion .h5."""
    import mgr_b200 as mgr
    from mgr_b200 import keras_io
    sp, sk, fu = _nets()
    p_sp, p_sk, p_fu = (str(tmp_path / n) for n in ("sp.h5", "sk.h5", "fu.h5"))
    for m, p in ((sp, p_sp), (sk, p_sk), (fu, p_fu)):
        keras_io.save_weights(m, p)
    sp2 = keras_io.model_from_json(keras_io.to_json(sp)); keras_io.load_weights(sp2, p_sp)
    sk2 = keras_io.model_from_json(keras_io.to_json(sk)); keras_io.load_weights(sk2, p_sk)
    fu2 = mgr.FusionNet(sp2, sk2, nb_classes=22, units=4, seed=99)
    keras_io.load_weights(fu2, p_fu)
    for (n1, a), (n2, b) in zip([e for v in keras_io.weight_table(fu).values() for e in v],
                                [e for v in keras_io.weight_table(fu2).values() for e in v]):
        assert n1 == n2 and np.array_equal(a, b)
    with pytest.raises(ValueError):
        keras_io.load_weights(sk2, p_sp)


def test_full_model_save_layout_is_found(tmp_path):
    """`model.save()` keeps the same tree under 'model_weights'."""
    from mgr_b200 import keras_h5
    w = keras_h5._Writer()
    k = np.arange(6, dtype=np.float32).reshape(2, 3)
    d = w.dataset(k)
    lg = w.group({"kernel:0": d}, {"weight_names": np.array([b"kernel:0"])})[0]
    mw = w.group({"dense_1": lg}, {"layer_names": np.array([b"dense_1"])})[0]
    root, t, h = w.group({"model_weights": mw}, {"keras_version": np.array(b"2.1.4")})
    path = str(tmp_path / "full.h5")
    open(path, "wb").write(w.finish(root, t, h))
    (lname, ws), = keras_h5.read_keras_weights(path)
    assert lname == "dense_1" and ws[0][0] == "kernel:0" and np.array_equal(ws[0][1], k)


def _msg(mtype, data):
    data = bytes(data) + b"\0" * (-len(data) % 8)
    return struct.pack("<HHB3x", mtype, len(data), 0) + data


def test_reader_variants_the_writer_does_not_emit(tmp_path):
    from mgr_b200 import keras_h5
    w = keras_h5._Writer()
    dt_f32 = w._dtype_msg(np.float32)
    dt_s5 = w._dtype_msg(np.dtype("S5"))
    # dataspace version 2 (no reserved bytes, explicit type), rank 1, dim 3
    ds_v2 = struct.pack("<BBBB", 2, 1, 0, 1) + struct.pack("<Q", 3)
    vals = np.array([1.5, -2.0, 3.25], dtype=np.float32)
    # attribute message version 3 (no padding, name encoding byte) and version 2
    name3 = b"alpha\0"
    attr_v3 = struct.pack("<BBHHHB", 3, 0, len(name3), len(dt_f32), len(ds_v2), 0) + name3 + dt_f32 + ds_v2 + vals.tobytes()
    name2 = b"tags\0"
    tags = np.array([b"ab", b"cdefg", b"h"], dtype="S5")
    attr_v2 = struct.pack("<BBHHH", 2, 0, len(name2), len(dt_s5), len(ds_v2)) + name2 + dt_s5 + ds_v2 + tags.tobytes()
    # compact dataset: data inside the layout message (version 3, class 0)
    small = np.array([[7, 8], [9, 10]], dtype=np.float32)
    compact = struct.pack("<BBH", 3, 0, small.nbytes) + small.tobytes()
    dset_msgs = [_msg(0x0001, w._dspace_msg(small.shape)), _msg(0x0003, dt_f32), _msg(0x0008, compact)]
    dset = w._header(dset_msgs)
    # group object header split over a continuation block: [symbol table, continuation] + [attr v3, attr v2]
    members = {"c": dset}
    _, tree_addr, heap_addr = w.group(members)
    cont = _msg(0x000C, attr_v3) + _msg(0x000C, attr_v2)
    cont_addr = w._alloc(cont)
    first = _msg(0x0011, struct.pack("<QQ", tree_addr, heap_addr)) + _msg(0x0010, struct.pack("<QQ", cont_addr, len(cont)))
    root = w._alloc(struct.pack("<BxHII4x", 1, 4, 1, len(first)) + first)
    path = str(tmp_path / "variants.h5")
    open(path, "wb").write(w.finish(root, tree_addr, heap_addr))
    f = keras_h5.H5Reader(path)
    a = f.attrs(f.root)
    assert np.array_equal(a["alpha"], vals)
    assert [t.rstrip(b"\0") for t in a["tags"].tolist()] == [b"ab", b"cdefg", b"h"]
    assert np.array_equal(f.dataset(f.resolve(f.root, "c")), small)


def test_reader_refuses_what_it_does_not_implement(tmp_path):
    from mgr_b200 import keras_h5
    with pytest.raises(keras_h5.H5FormatError):
        keras_h5.H5Reader(b"not an hdf5 file at all......")
    sp, _, _ = _nets()
    from mgr_b200 import keras_io
    path = str(tmp_path / "w.h5")
    keras_io.save_weights(sp, path)
    raw = bytearray(open(path, "rb").read())
    raw[8] = 2                                          # superblock version 2 = libver 'latest'
    with pytest.raises(keras_h5.H5FormatError):
        keras_h5.H5Reader(bytes(raw))
    # a filter pipeline message (compressed dataset) must be rejected, not mis-read
    w = keras_h5._Writer()
    arr = np.ones((2, 2), np.float32)
    data = w._alloc(arr.tobytes())
    msgs = [_msg(0x0001, w._dspace_msg(arr.shape)), _msg(0x0003, w._dtype_msg(arr.dtype)),
            _msg(0x0008, struct.pack("<BBQQ", 3, 1, data, arr.nbytes)), _msg(0x000B, b"\1\1" + b"\0" * 6)]
    d = w._header(msgs)
    root, t, h = w.group({"z": d})
    f = keras_h5.H5Reader(w.finish(root, t, h))
    with pytest.raises(keras_h5.H5FormatError):
        f.dataset(f.resolve(f.root, "z"))
