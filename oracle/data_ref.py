"""Oracle: the reference's batch generators restated literally (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows /root/reference/audio_network/data_generator.py:52-268 and
/root/reference/multimodal_fusion/data_generator.py:68-300 statement by statement: the CSVs are re-read
for every batch, arrays are NumPy float64 initialised with ones/zeros, `keras.preprocessing.sequence.
pad_sequences` (Keras 2.1.4, not vendored) is restated for the one call shape the reference uses
(`padding='post', truncating='post'`), `sklearn.preprocessing.scale` as (x - mean) / population-std with
zero-variance columns left unscaled, and Python 2's `random.shuffle` (the reference's interpreter) as
`j = int(random() * (i + 1))` walking i downwards.  Parity unpinned by reference fixtures (the reference
ships no data); pinned by hand-checked tiny datasets in tests/test_data_generator.py.
"""
import os
import random
import re

import numpy as np
import pandas as pd

CLASS_2_WORDS = {0: [0.0], 1: [1.0], 2: [2.0, 3.0], 3: [4.0], 4: [5.0, 6.0, 7.0], 5: [8.0, 9.0, 10.0], 6: [8.0, 11.0],
                 7: [12.0, 13.0], 8: [14.0, 15.0], 9: [16.0, 17.0], 10: [18.0, 19.0, 20.0, 21.0, 22.0], 11: [23.0],
                 12: [24.0, 25.0, 26.0], 13: [27.0], 14: [28.0, 11.0, 29.0], 15: [18.0, 30.0, 31.0, 32.0],
                 16: [33.0, 34.0], 17: [35.0, 36.0, 37.0], 18: [38.0], 19: [39.0, 40.0, 41.0, 13.0], 20: [40.0, 42.0],
                 21: [43.0]}
SKEL = ['lh_v', 'rh_v', 'le_v', 're_v', 'lh_dist_rp', 'rh_dist_rp', 'lh_hip_d', 'rh_hip_d', 'le_hip_d', 're_hip_d',
        'lh_shc_d', 'rh_shc_d', 'le_shc_d', 're_shc_d', 'lh_hip_ang', 'rh_hip_ang', 'lh_shc_ang', 'rh_shc_ang',
        'lh_el_ang', 'rh_el_ang']


class _Py2Random:
    """The global `random` module of the reference process: seed(10), then shuffle() calls in program order."""

    def __init__(self):
        self.r = random.Random()

    def seed(self, a):
        self.r.seed(a)

    def shuffle(self, x):
        for i in reversed(range(1, len(x))):
            j = int(self.r.random() * (i + 1))
            x[i], x[j] = x[j], x[i]


def pad_sequences_post(seqs, maxlen, dtype='int32', value=0.):
    """keras.preprocessing.sequence.pad_sequences(seqs, maxlen, padding='post', truncating='post', dtype, value)."""
    sample_shape = tuple()
    for s in seqs:
        if len(s) > 0:
            sample_shape = np.asarray(s).shape[1:]
            break
    x = (np.ones((len(seqs), maxlen) + sample_shape) * value).astype(dtype)
    for idx, s in enumerate(seqs):
        if not len(s):
            continue
        trunc = np.asarray(s[:maxlen], dtype=dtype)
        x[idx, :len(trunc)] = trunc
    return x


def sent_2_words(lab_seq):
    new_seq = []
    for lab in lab_seq:
        new_seq = new_seq + CLASS_2_WORDS[lab]
    return np.asarray(new_seq)


class AudioGeneratorRef:
    def __init__(self, minibatch_size, numfeats, maxlen, nb_classes, dataset, val_split=0.2,
                 absolute_max_sequence_len=150, data_root='../data'):
        self.minibatch_size, self.maxlen, self.numfeats = minibatch_size, maxlen, numfeats
        self.val_split, self.absolute_max_sequence_len = val_split, absolute_max_sequence_len
        self.train_index = self.val_index = 0
        self.nb_classes = nb_classes
        self.blank_label = np.array([nb_classes - 1])
        self.dataset = dataset
        self.random = _Py2Random()
        self.in_dir = os.path.join(data_root, 'train_audio' if dataset == 'train' else 'val_audio')
        lab_file = os.path.join(data_root, 'training_oov.csv' if dataset == 'train' else 'validation.csv')
        self.labs = pd.read_csv(lab_file)
        file_list = sorted([int(re.findall(r'audio_(\d+).csv', n)[0]) for n in os.listdir(self.in_dir)])
        _split(self, file_list)

    def get_batch(self, train):
        file_list, index = (self.train_list, self.train_index) if train else (self.val_list, self.val_index)
        batch = file_list[index:(index + self.minibatch_size)]
        size = len(batch)
        X_data = np.ones([size, self.maxlen, self.numfeats])
        labels = np.ones([size, self.absolute_max_sequence_len])
        input_length = np.zeros([size, 1])
        label_length = np.zeros([size, 1])
        for i in range(len(batch)):
            file = batch[i]
            vf = pd.read_csv(os.path.join(self.in_dir, 'audio_' + str(file) + '.csv')).drop(['file_number'], axis=1)
            if set(['39', '40']).issubset(vf.columns):
                vf = vf.drop(['39', '40'], axis=1)
            vf = vf.iloc[::5, :].reset_index(drop=True)
            gest_seq = vf.values.astype(float)
            gest_seq = pad_sequences_post([gest_seq], maxlen=self.maxlen, dtype='float32')
            lab_seq = self.labs[self.labs['Id'] == file]
            lab_seq = np.array([int(lab) for lab in lab_seq['Sequence'].values[0].split()]).astype('float32')
            lab_seq = sent_2_words(lab_seq)
            if lab_seq.shape[0] == 0:
                lab_seq = pad_sequences_post([self.blank_label], maxlen=self.absolute_max_sequence_len, value=-1)
                labels[i, :] = lab_seq
                label_length[i] = 1
            else:
                X_data[i, :, :] = gest_seq
                label_length[i] = lab_seq.shape[0]
                lab_seq = pad_sequences_post([lab_seq], maxlen=self.absolute_max_sequence_len, value=-1)
                labels[i, :] = lab_seq
            input_length[i] = (X_data[i].shape[0] - 2)
        inputs = {'the_input': X_data, 'the_labels': labels, 'input_length': input_length,
                  'label_length': label_length}
        return inputs, {'ctc': np.zeros([size])}


class FusionGeneratorRef:
    def __init__(self, minibatch_size, numfeats_skeletal, numfeats_speech, maxlen, nb_classes, dataset,
                 val_split=0.2, absolute_max_sequence_len=35, data_root='../data'):
        self.minibatch_size, self.maxlen = minibatch_size, maxlen
        self.numfeats_speech, self.numfeats_skeletal = numfeats_speech, numfeats_skeletal
        self.val_split, self.absolute_max_sequence_len = val_split, absolute_max_sequence_len
        self.train_index = self.val_index = 0
        self.nb_classes = nb_classes
        self.blank_label = np.array([nb_classes - 1])
        self.dataset = dataset
        self.random = _Py2Random()
        names = {'train': ('train_audio', 'Training_set_skeletal.csv', 'training_oov.csv'),
                 'val': ('val_audio', 'Validation_set_skeletal.csv', 'validation.csv'),
                 'final': ('final_audio', 'final_set_skeletal.csv', 'validation.csv')}[dataset]
        self.in_audio_dir = os.path.join(data_root, names[0])
        self.labs = pd.read_csv(os.path.join(data_root, names[2]))
        self.df_s = pd.read_csv(os.path.join(data_root, names[1]))
        data = self.df_s[SKEL].values.astype(float)
        mean, std = data.mean(axis=0), data.std(axis=0)
        std[std == 0.0] = 1.0
        norm_df = pd.DataFrame((data - mean) / std, columns=SKEL)
        norm_df['file_number'] = self.df_s['file_number']
        self.df_s = norm_df
        file_list = sorted([int(re.findall(r'audio_(\d+).csv', n)[0]) for n in os.listdir(self.in_audio_dir)])
        _split(self, file_list)

    def get_batch(self, train):
        file_list, index = (self.train_list, self.train_index) if train else (self.val_list, self.val_index)
        batch = file_list[index:(index + self.minibatch_size)]
        size = len(batch)
        X_data_a = np.ones([size, self.maxlen, self.numfeats_speech])
        labels_a = np.ones([size, self.absolute_max_sequence_len])
        input_length_a = np.zeros([size, 1])
        label_length_a = np.zeros([size, 1])
        X_data_s = np.ones([size, self.maxlen, self.numfeats_skeletal])
        for i in range(len(batch)):
            file = batch[i]
            vf_a = pd.read_csv(os.path.join(self.in_audio_dir, 'audio_' + str(file) + '.csv')).drop(['file_number'], axis=1)
            if set(['39', '40']).issubset(vf_a.columns):
                vf_a = vf_a.drop(['39', '40'], axis=1)
            vf_a = vf_a.iloc[::5, :].reset_index(drop=True)
            vf_s = self.df_s[self.df_s['file_number'] == file]
            gest_seq_a = pad_sequences_post([vf_a.values.astype(float)], maxlen=self.maxlen, dtype='float32')
            gest_seq_s = pad_sequences_post([vf_s[SKEL].values.astype(float)], maxlen=self.maxlen, dtype='float32')
            if self.dataset != 'final':
                lab_seq = self.labs[self.labs['Id'] == file]
                lab_seq = np.array([int(lab) for lab in lab_seq['Sequence'].values[0].split()]).astype('float32')
            else:
                lab_seq = np.array([0])
            if lab_seq.shape[0] == 0:
                lab_seq = pad_sequences_post([self.blank_label], maxlen=self.absolute_max_sequence_len, value=-1)
                labels_a[i, :] = lab_seq
                label_length_a[i] = 1
            else:
                X_data_a[i, :, :] = gest_seq_a
                try:
                    X_data_s[i, :, :] = gest_seq_s
                except Exception:
                    pass   # print 'blank'
                label_length_a[i] = lab_seq.shape[0]
                lab_seq = pad_sequences_post([lab_seq], maxlen=self.absolute_max_sequence_len, value=-1)
                labels_a[i, :] = lab_seq
            input_length_a[i] = (X_data_a[i].shape[0] - 2)
        inputs = {'the_input_audio': X_data_a, 'the_input_skeletal': X_data_s, 'the_labels': labels_a,
                  'input_length': input_length_a, 'label_length': label_length_a}
        return inputs, {'ctc': np.zeros([size])}


def _split(self, file_list):
    if self.dataset == 'train':
        self.random.seed(10)
        self.random.shuffle(file_list)
        split_point = int(len(file_list) * (1 - self.val_split))
        self.train_list, self.val_list = file_list[:split_point], file_list[split_point:]
        self.train_size = len(self.train_list)
        self.val_size = len(self.val_list)
        train_mod_by_batch_size = self.train_size % self.minibatch_size
        if train_mod_by_batch_size != 0:
            del self.train_list[-train_mod_by_batch_size:]
            self.train_size -= train_mod_by_batch_size
        val_mod_by_batch_size = self.val_size % self.minibatch_size
        if val_mod_by_batch_size != 0:
            del self.val_list[-val_mod_by_batch_size:]
            self.val_size -= val_mod_by_batch_size
    else:
        self.random.seed(10)
        self.train_list, self.train_size = [], 0
        self.val_list = file_list
        self.val_size = len(self.val_list)


def next_batches(gen, train, n):
    """What `next_train()` / `next_val()` yield for n draws (index bookkeeping of data_generator.py:243-268)."""
    out = []
    for _ in range(n):
        out.append(gen.get_batch(train))
        if train:
            gen.train_index += gen.minibatch_size
            if gen.train_index >= gen.train_size:
                gen.train_index = 0
        else:
            gen.val_index += gen.minibatch_size
            if gen.val_index >= gen.val_size:
                gen.val_index = 0
    return out


def on_epoch_end(gen):
    gen.train_index = 0
    gen.val_index = 0
    gen.random.shuffle(gen.train_list)
    gen.random.shuffle(gen.val_list)
