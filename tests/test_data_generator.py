"""Host-side batch producers against the literal restatement of the reference's
generators in oracle/data_ref.py, on small synthetic datasets written to a temp directory (CPU only)."""
import csv
import os

import numpy as np
import pytest


def _write_dataset(root, kind, ids, rng, blank_ids=(), no_skeletal=(), long_ids=(), extra_cols=True):
    audio_dir = {"train": "train_audio", "val": "val_audio", "final": "final_audio"}[kind]
    os.makedirs(os.path.join(root, audio_dir), exist_ok=True)
    for fid in ids:
        n = int(rng.integers(20, 90)) if fid not in long_ids else 400
        cols = [str(c) for c in range(39)] + (["39", "40"] if extra_cols and fid % 2 == 0 else []) + ["file_number"]
        with open(os.path.join(root, audio_dir, "audio_%d.csv" % fid), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(cols)
            for _ in range(n):
                row = list(np.round(rng.standard_normal(39), 5))
                if "39" in cols:
                    row += [fid, 3]
                w.writerow(row + [fid])
    from oracle.data_ref import SKEL
    skel_name = {"train": "Training_set_skeletal.csv", "val": "Validation_set_skeletal.csv",
                 "final": "final_set_skeletal.csv"}[kind]
    with open(os.path.join(root, skel_name), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["frame"] + SKEL + ["file_number", "label"])
        for fid in ids:
            if fid in no_skeletal:
                continue
            for fr in range(int(rng.integers(5, 40)) if fid not in long_ids else 80):
                vals = list(np.round(rng.standard_normal(len(SKEL)) * 3 + 1, 5))
                vals[7] = 2.5           # a constant column: z-scoring must leave it at zero, not NaN
                w.writerow([fr] + vals + [fid, 0])
    lab_name = "training_oov.csv" if kind == "train" else "validation.csv"
    with open(os.path.join(root, lab_name), "w", newline="") as f:
        w = csv.writer(f, quoting=csv.QUOTE_ALL)
        w.writerow(["Id", "Sequence"])
        for fid in ids:
            seq = " " if fid in blank_ids else " ".join(str(int(v)) for v in rng.integers(1, 21, size=int(rng.integers(1, 9))))
            w.writerow([fid, seq])


def _same(a, b):
    assert a.keys() == b.keys()
    for k in a:
        assert a[k].shape == b[k].shape, k
        assert np.array_equal(np.asarray(a[k], dtype=np.float64), np.asarray(b[k], dtype=np.float64)), k


def test_python2_shuffle_stream():
    """seed(10) gives the Mersenne-Twister stream both interpreters share; the Python-2 shuffle consumes one
    draw per position, from the back (known first draws: 0.5714025946899135, 0.4288890546751146)."""
    import random
    from mgr_b200.data_generator import py2_shuffle
    r = random.Random(10)
    assert (r.random(), r.random()) == (0.5714025946899135, 0.4288890546751146)
    x = list(range(5))
    py2_shuffle(x, random.Random(10))
    # i=4: j=int(.5714*5)=2 -> [0,1,4,3,2]; i=3: j=int(.4289*4)=1 -> [0,3,4,1,2]; i=2: j=int(.5781*3)=1 -> [0,4,3,1,2];
    # i=1: j=int(.2061*2)=0 -> [4,0,3,1,2]
    assert x == [4, 0, 3, 1, 2]


def test_sent_2_words_and_padding():
    from mgr_b200.data_generator import sent_2_words, pad_post
    from oracle import data_ref
    seq = np.array([2, 10, 21, 0], dtype=np.float32)
    assert np.array_equal(sent_2_words(seq), data_ref.sent_2_words(seq))
    assert sent_2_words(np.array([])).shape == (0,)
    a = np.arange(12, dtype=float).reshape(6, 2)
    assert np.array_equal(pad_post(a, 4), data_ref.pad_sequences_post([a], 4, dtype="float32")[0])
    assert np.array_equal(pad_post(a, 9), data_ref.pad_sequences_post([a], 9, dtype="float32")[0])
    assert np.array_equal(pad_post(np.array([21]), 5, value=-1), data_ref.pad_sequences_post([np.array([21])], 5, value=-1)[0])


@pytest.mark.parametrize("kind", ["train", "val"])
def test_audio_generator_matches_reference_restatement(tmp_path, kind):
    from mgr_b200.data_generator import AudioDataGenerator
    from oracle import data_ref
    rng = np.random.default_rng(1)
    ids = list(range(3, 26))
    _write_dataset(str(tmp_path), kind, ids, rng, blank_ids=(5, 17), long_ids=(9,))
    args = dict(minibatch_size=4, numfeats=39, maxlen=30, nb_classes=44, dataset=kind, val_split=0.25,
                absolute_max_sequence_len=40, data_root=str(tmp_path))
    g, r = AudioDataGenerator(**args), data_ref.AudioGeneratorRef(**args)
    assert g.get_file_list(True) == r.train_list and g.get_file_list(False) == r.val_list
    assert (g.get_size(True), g.get_size(False)) == (r.train_size, r.val_size)
    for epoch in range(2):
        for train in ([True, False] if kind == "train" else [False]):
            n = (g.get_size(train) // 4) + 2          # wraps around the end of the list
            it = g.next_train() if train else g.next_val()
            for (gi, go), (ri, ro) in zip((next(it) for _ in range(n)), data_ref.next_batches(r, train, n)):
                _same(gi, ri)
                assert np.array_equal(go["ctc"], ro["ctc"])
                assert gi["the_input"].dtype == np.float32 and gi["input_length"].dtype == np.int64
        g.on_epoch_end()
        data_ref.on_epoch_end(r)
        assert g.get_file_list(True) == r.train_list and g.get_file_list(False) == r.val_list


@pytest.mark.parametrize("kind", ["train", "final"])
def test_fusion_generator_matches_reference_restatement(tmp_path, kind):
    from mgr_b200.data_generator import FusionDataGenerator
    from oracle import data_ref
    rng = np.random.default_rng(2)
    ids = list(range(100, 121))
    _write_dataset(str(tmp_path), kind, ids, rng, blank_ids=(104,), no_skeletal=(110,), long_ids=(101,))
    args = dict(minibatch_size=2, numfeats_skeletal=20, numfeats_speech=39, maxlen=25, nb_classes=22, dataset=kind,
                val_split=0.2, absolute_max_sequence_len=35, data_root=str(tmp_path))
    g, r = FusionDataGenerator(**args), data_ref.FusionGeneratorRef(**args)
    assert g.get_file_list(True) == r.train_list and g.get_file_list(False) == r.val_list
    for train in ([True, False] if kind == "train" else [False]):
        n = g.get_size(train) // 2 + 1
        it = g.next_train() if train else g.next_val()
        for (gi, _), (ri, _) in zip((next(it) for _ in range(n)), data_ref.next_batches(r, train, n)):
            _same(gi, ri)
            assert np.all(gi["input_length"] == 23)
    # the example without labels keeps the all-ones data and the blank label (train set only: 'final' has no labels)
    if kind == "train":
        g.train_list, g.train_index = [104, 110], 0
        inp, _ = g.get_batch(True)
        assert np.all(inp["the_input_audio"][0] == 1) and np.all(inp["the_input_skeletal"][0] == 1)
        assert inp["the_labels"][0, 0] == 21 and np.all(inp["the_labels"][0, 1:] == -1) and inp["label_length"][0, 0] == 1
        assert np.all(inp["the_input_skeletal"][1] == 1) and not np.all(inp["the_input_audio"][1] == 1)


def test_to_device_argument_order(tmp_path):
    import torch
    from mgr_b200.data_generator import FusionDataGenerator, to_device
    rng = np.random.default_rng(3)
    _write_dataset(str(tmp_path), "val", list(range(1, 7)), rng)
    g = FusionDataGenerator(2, 20, 39, 20, 22, "val", data_root=str(tmp_path))
    inp, _ = next(g.next_val())
    xa, xs, lab, il, ll = to_device(inp, torch.device("cpu"), pinned=False)
    assert xa.shape == (2, 20, 39) and xs.shape == (2, 20, 20) and lab.shape == (2, 35)
    assert il.dtype == torch.int64 and ll.dtype == torch.int64 and xa.dtype == torch.float32


def test_generators_agree_on_random_configurations(tmp_path):
    """Random mini-batch sizes, split fractions, sequence lengths and dataset sizes (seeded sweep rather than
    hypothesis: each case writes a small dataset to disk)."""
    from mgr_b200.data_generator import AudioDataGenerator, FusionDataGenerator
    from oracle import data_ref
    rng = np.random.default_rng(11)
    for case in range(4):
        root = tmp_path / ("case%d" % case)
        os.makedirs(root)
        n_files = int(rng.integers(7, 19))
        ids = sorted(rng.choice(np.arange(1, 400), size=n_files, replace=False).tolist())
        blank = tuple(rng.choice(ids, size=2, replace=False).tolist())
        _write_dataset(str(root), "train", ids, rng, blank_ids=blank, no_skeletal=(ids[0],), long_ids=(ids[-1],))
        mb = int(rng.integers(1, 5))
        vs = float(rng.choice([0.1, 0.2, 0.34, 0.5]))
        maxlen = int(rng.integers(8, 40))
        a_args = dict(minibatch_size=mb, numfeats=39, maxlen=maxlen, nb_classes=44, dataset="train", val_split=vs,
                      absolute_max_sequence_len=60, data_root=str(root))
        f_args = dict(minibatch_size=mb, numfeats_skeletal=20, numfeats_speech=39, maxlen=maxlen, nb_classes=22,
                      dataset="train", val_split=vs, absolute_max_sequence_len=35, data_root=str(root))
        for g, r in ((AudioDataGenerator(**a_args), data_ref.AudioGeneratorRef(**a_args)),
                     (FusionDataGenerator(**f_args), data_ref.FusionGeneratorRef(**f_args))):
            assert g.get_file_list(True) == r.train_list and g.get_file_list(False) == r.val_list
            for train in (True, False):
                if g.get_size(train) == 0:
                    continue
                n = g.get_size(train) // mb + 1
                it = g.next_train() if train else g.next_val()
                for (gi, _), (ri, _) in zip((next(it) for _ in range(n)), data_ref.next_batches(r, train, n)):
                    _same(gi, ri)
            g.on_epoch_end()
            data_ref.on_epoch_end(r)
            assert g.get_file_list(True) == r.train_list and g.get_file_list(False) == r.val_list
