"""Host-side batch producers with the reference's `DataGenerator` interface and batch format.

`AudioDataGenerator` mirrors /root/reference/audio_network/data_generator.py:18-268 (speech network:
`audio_<id>.csv` MFCC files, every 5th frame, word-level labels through `sent_2_words`);
`FusionDataGenerator` mirrors /root/reference/multimodal_fusion/data_generator.py:20-300 (audio files
+ one skeletal CSV, z-scored once, gesture-level labels).  Same constructor arguments, the same
`get_batch / next_train / next_val / get_size / get_file_list / on_epoch_end` methods, the same
`(inputs, outputs)` dictionaries -- with the dtypes Keras feeds the graph with (float32 data and labels,
int64 lengths) instead of NumPy's default float64.

What is different is how a batch is produced: every file is parsed ONCE (the reference re-reads and
re-filters the CSVs for every batch of every epoch; the skeletal table is grouped by file once instead of
one boolean filter per example), sequences are kept already down-sampled and padded, and `get_batch`
is a gather of cached rows, so that the loader keeps up with a GPU step that takes tens of milliseconds.
`to_device()` hands a batch (through pinned memory) to `FusionTrainer.step` / `ctc_lambda_func` in their
argument order.

Reference behaviours kept on purpose (the oracle in oracle/data_ref.py restates them literally):
  * train/validation split: Python-2 `random.seed(10); random.shuffle(file_list)` (reproduced here with the
    Python-2 shuffle algorithm on the same Mersenne-Twister stream), tail trimmed to whole mini-batches;
  * rows without labels become the "blank example": label `[nb_classes-1]`, label_length 1, and the DATA STAYS
    ALL ONES (`np.ones` initialisation, data_generator.py:173,207-213);
  * `input_length` is always `maxlen - 2` (`:223`), whatever the true length -- padding frames are zeros;
  * labels padded with -1 to `absolute_max_sequence_len`.
"""
import os
import random
import re

import numpy as np
import pandas as pd

# gesture class -> word ids of the 44-word speech vocabulary (audio_network/data_generator.py:138-141)
_CLASS_WORDS = ((0,), (1,), (2, 3), (4,), (5, 6, 7), (8, 9, 10), (8, 11), (12, 13), (14, 15), (16, 17),
                (18, 19, 20, 21, 22), (23,), (24, 25, 26), (27,), (28, 11, 29), (18, 30, 31, 32), (33, 34),
                (35, 36, 37), (38,), (39, 40, 41, 13), (40, 42), (43,))
SKELETAL_COLUMNS = ['lh_v', 'rh_v', 'le_v', 're_v', 'lh_dist_rp', 'rh_dist_rp', 'lh_hip_d', 'rh_hip_d', 'le_hip_d',
                    're_hip_d', 'lh_shc_d', 'rh_shc_d', 'le_shc_d', 're_shc_d', 'lh_hip_ang', 'rh_hip_ang',
                    'lh_shc_ang', 'rh_shc_ang', 'lh_el_ang', 'rh_el_ang']


def sent_2_words(lab_seq):
    """Gesture-level label sequence -> word-level label sequence (float array), data_generator.py:121-149."""
    out = []
    for lab in lab_seq:
        out.extend(float(w) for w in _CLASS_WORDS[int(lab)])
    return np.asarray(out)


def py2_shuffle(x, rng):
    """`random.shuffle` as Python 2 does it (the reference's interpreter): j = int(random() * (i + 1))."""
    for i in reversed(range(1, len(x))):
        j = int(rng.random() * (i + 1))
        x[i], x[j] = x[j], x[i]


def pad_post(seq, maxlen, width=None, value=0.0, dtype=np.float32):
    """keras `pad_sequences([seq], maxlen, padding='post', truncating='post')[0]` for one 1-D / 2-D sequence."""
    seq = np.asarray(seq)
    shape = (maxlen,) + ((seq.shape[1],) if seq.ndim == 2 else ((width,) if width else ()))
    out = np.full(shape, value, dtype=dtype)
    n = min(maxlen, seq.shape[0])
    out[:n] = seq[:n]
    return out


def _label_sequence(labs, file_id):
    row = labs[labs['Id'] == file_id]
    return np.array([int(v) for v in row['Sequence'].values[0].split()]).astype('float32')


class _GeneratorBase:
    """Split, indexing and epoch logic shared by the two generators (identical in both reference files)."""

    def _split(self, file_list):
        self._rng = random.Random(10)          # `random.seed(10)`: one stream for the split and every epoch's reshuffle
        if self.dataset == 'train':
            py2_shuffle(file_list, self._rng)
            split_point = int(len(file_list) * (1 - self.val_split))
            self.train_list, self.val_list = file_list[:split_point], file_list[split_point:]
            for name in ('train_list', 'val_list'):
                lst = getattr(self, name)
                extra = len(lst) % self.minibatch_size
                if extra:
                    del lst[-extra:]
            self.train_size, self.val_size = len(self.train_list), len(self.val_list)
        else:
            self.train_list, self.train_size = [], 0
            self.val_list = file_list
            self.val_size = len(file_list)

    def get_size(self, train):
        return self.train_size if train else self.val_size

    def get_file_list(self, train):
        return self.train_list if train else self.val_list

    def _batch_ids(self, train):
        lst, idx = (self.train_list, self.train_index) if train else (self.val_list, self.val_index)
        return lst[idx:idx + self.minibatch_size]

    def next_train(self):
        while True:
            ret = self.get_batch(train=True)
            self.train_index += self.minibatch_size
            if self.train_index >= self.train_size:
                self.train_index = 0
            yield ret

    def next_val(self):
        while True:
            ret = self.get_batch(train=False)
            self.val_index += self.minibatch_size
            if self.val_index >= self.val_size:
                self.val_index = 0
            yield ret

    def on_epoch_end(self, epoch=None, logs=None):
        """Reset the cursors and reshuffle both lists (the reference also saves the model here; that is the
        caller's business in this package)."""
        self.train_index = 0
        self.val_index = 0
        py2_shuffle(self.train_list, self._rng)
        py2_shuffle(self.val_list, self._rng)

    def _audio_sequence(self, path):
        vf = pd.read_csv(path).drop(['file_number'], axis=1)
        if {'39', '40'}.issubset(vf.columns):
            vf = vf.drop(['39', '40'], axis=1)
        return vf.iloc[::5, :].to_numpy().astype(float)     # every 5th MFCC frame


class AudioDataGenerator(_GeneratorBase):
    """audio_network/data_generator.py:18 `DataGenerator` (speech BLSTM-CTC on word labels)."""

    def __init__(self, minibatch_size, numfeats, maxlen, nb_classes, dataset, val_split=0.2,
                 absolute_max_sequence_len=150, data_root='../data'):
        self.minibatch_size, self.maxlen, self.numfeats = minibatch_size, maxlen, numfeats
        self.val_split, self.absolute_max_sequence_len = val_split, absolute_max_sequence_len
        self.train_index = self.val_index = 0
        self.nb_classes = nb_classes
        self.blank_label = np.array([nb_classes - 1])
        self.dataset = dataset
        self.in_dir = os.path.join(data_root, {'train': 'train_audio', 'val': 'val_audio'}[dataset])
        self.lab_file = os.path.join(data_root, {'train': 'training_oov.csv', 'val': 'validation.csv'}[dataset])
        self.build_dataset()

    def build_dataset(self):
        self.labs = pd.read_csv(self.lab_file)
        ids = sorted(int(re.findall(r'audio_(\d+).csv', f)[0]) for f in os.listdir(self.in_dir))
        self._split(ids)
        self._cache = {}

    def _example(self, file_id):
        ex = self._cache.get(file_id)
        if ex is None:
            x = pad_post(self._audio_sequence(os.path.join(self.in_dir, 'audio_%d.csv' % file_id)), self.maxlen,
                         width=self.numfeats)
            lab = sent_2_words(_label_sequence(self.labs, file_id))
            ex = self._cache[file_id] = self._pack(x, lab)
        return ex

    def _pack(self, x, lab):
        if lab.shape[0] == 0:     # "blank example": data stays all ones
            return (np.ones((self.maxlen, x.shape[1]), np.float32),
                    pad_post(self.blank_label, self.absolute_max_sequence_len, value=-1), 1)
        return x, pad_post(lab, self.absolute_max_sequence_len, value=-1), lab.shape[0]

    def get_batch(self, train):
        batch = self._batch_ids(train)
        size = len(batch)
        X = np.ones((size, self.maxlen, self.numfeats), np.float32)
        labels = np.ones((size, self.absolute_max_sequence_len), np.float32)
        label_length = np.zeros((size, 1), np.int64)
        for i, fid in enumerate(batch):
            x, lab, n = self._example(fid)
            X[i], labels[i], label_length[i] = x, lab, n
        input_length = np.full((size, 1), self.maxlen - 2, np.int64)
        inputs = {'the_input': X, 'the_labels': labels, 'input_length': input_length, 'label_length': label_length}
        return inputs, {'ctc': np.zeros([size])}


class FusionDataGenerator(_GeneratorBase):
    """multimodal_fusion/data_generator.py:20 `DataGenerator` (speech + skeletal inputs, gesture labels)."""

    def __init__(self, minibatch_size, numfeats_skeletal, numfeats_speech, maxlen, nb_classes, dataset,
                 val_split=0.2, absolute_max_sequence_len=35, data_root='../data'):
        self.minibatch_size, self.maxlen = minibatch_size, maxlen
        self.numfeats_speech, self.numfeats_skeletal = numfeats_speech, numfeats_skeletal
        self.val_split, self.absolute_max_sequence_len = val_split, absolute_max_sequence_len
        self.train_index = self.val_index = 0
        self.nb_classes = nb_classes
        self.blank_label = np.array([nb_classes - 1])
        self.dataset = dataset
        sub = {'train': ('train_audio', 'Training_set_skeletal.csv', 'training_oov.csv'),
               'val': ('val_audio', 'Validation_set_skeletal.csv', 'validation.csv'),
               'final': ('final_audio', 'final_set_skeletal.csv', 'validation.csv')}[dataset]
        self.in_audio_dir = os.path.join(data_root, sub[0])
        self.in_file_skeletal = os.path.join(data_root, sub[1])
        self.lab_file = os.path.join(data_root, sub[2])
        self.load_dataset()

    def load_dataset(self):
        self.labs = pd.read_csv(self.lab_file)
        self.df_s = pd.read_csv(self.in_file_skeletal)
        self.df_s = self.normalize_data()
        # one pass over the skeletal table instead of one boolean filter per example per batch
        self._skel = {int(k): g[SKELETAL_COLUMNS].to_numpy().astype(float)
                      for k, g in self.df_s.groupby('file_number', sort=False)}
        ids = sorted(int(re.findall(r'audio_(\d+).csv', f)[0]) for f in os.listdir(self.in_audio_dir))
        self._split(ids)
        self._cache = {}

    def normalize_data(self):
        """Zero mean / unit variance per column over the whole table (`sklearn.preprocessing.scale`:
        population standard deviation, constant columns left at zero)."""
        data = self.df_s[SKELETAL_COLUMNS].to_numpy().astype(float)
        mean = data.mean(axis=0)
        std = data.std(axis=0)
        std[std == 0.0] = 1.0
        norm_df = pd.DataFrame((data - mean) / std, columns=SKELETAL_COLUMNS)
        norm_df['file_number'] = self.df_s['file_number'].to_numpy()
        return norm_df

    def _example(self, file_id):
        ex = self._cache.get(file_id)
        if ex is None:
            xa = pad_post(self._audio_sequence(os.path.join(self.in_audio_dir, 'audio_%d.csv' % file_id)),
                          self.maxlen, width=self.numfeats_speech)
            sk = self._skel.get(file_id)
            # a file without skeletal frames keeps the all-ones initialisation (the reference's `except: print 'blank'`)
            xs = (np.ones((self.maxlen, self.numfeats_skeletal), np.float32) if sk is None or sk.shape[0] == 0
                  else pad_post(sk, self.maxlen, width=self.numfeats_skeletal))
            lab = _label_sequence(self.labs, file_id) if self.dataset != 'final' else np.array([0])
            if lab.shape[0] == 0:
                ex = (np.ones_like(xa), np.ones_like(xs),
                      pad_post(self.blank_label, self.absolute_max_sequence_len, value=-1), 1)
            else:
                ex = (xa, xs, pad_post(lab, self.absolute_max_sequence_len, value=-1), lab.shape[0])
            self._cache[file_id] = ex
        return ex

    def get_batch(self, train):
        batch = self._batch_ids(train)
        size = len(batch)
        Xa = np.ones((size, self.maxlen, self.numfeats_speech), np.float32)
        Xs = np.ones((size, self.maxlen, self.numfeats_skeletal), np.float32)
        labels = np.ones((size, self.absolute_max_sequence_len), np.float32)
        label_length = np.zeros((size, 1), np.int64)
        for i, fid in enumerate(batch):
            xa, xs, lab, n = self._example(fid)
            Xa[i], Xs[i], labels[i], label_length[i] = xa, xs, lab, n
        input_length = np.full((size, 1), self.maxlen - 2, np.int64)
        inputs = {'the_input_audio': Xa, 'the_input_skeletal': Xs, 'the_labels': labels,
                  'input_length': input_length, 'label_length': label_length}
        return inputs, {'ctc': np.zeros([size])}


def to_device(inputs, device, pinned=True):
    """A generator batch -> the argument tuple of `FusionTrainer.step` (fusion batches: xa, xs, labels,
    input_length, label_length) or of `ctc_lambda_func` callers (audio batches: x, labels, input_length,
    label_length), copied from pinned host memory with non-blocking copies."""
    import torch
    keys = (['the_input_audio', 'the_input_skeletal'] if 'the_input_audio' in inputs else ['the_input'])
    keys += ['the_labels', 'input_length', 'label_length']
    out = []
    for k in keys:
        t = torch.from_numpy(np.ascontiguousarray(inputs[k]))
        if pinned and torch.cuda.is_available():
            t = t.pin_memory()
        out.append(t.to(device, non_blocking=True))
    return tuple(out)
