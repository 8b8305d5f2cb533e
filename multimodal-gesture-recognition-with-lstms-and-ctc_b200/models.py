"""The three reference topologies on the B200 kernels, plus the fusion training step.

  * `SpeechNet`  -- audio_network/speech_lstm_ctc_words.py:32-134: GaussianNoise(0.5) ->
    BLSTM(500, dropout .4) -> BLSTM(500, dropout .5) -> add -> Dropout(.5) -> Dense(44) -> softmax.
  * `SkeletalNet` -- skeletal_network/skeletal_lstm_ctc.py:296-394: GaussianNoise(0.5) ->
    BLSTM(300, .6) -> BLSTM(300, .6) -> add -> Dropout(.6) -> Dense(22) -> softmax.
  * `FusionNet` -- multimodal_fusion/multimodal.py:58-215: the two towers (their BLSTM layers
    re-used and frozen, :109-148) -> concat (speech first, :155-156) -> BLSTM(100, dropout .5) ->
    Dropout(.5) -> Dense(22) -> softmax -> ctc_lambda_func; Adam(1e-4, clipvalue .5, decay 1e-5).
All regularisers are explicit arguments (inject them for parity, or draw them with
`sample_regularisers`, on-device Philox) because K.set_learning_phase(1) keeps them active in
training AND in the frozen towers (speech:40, multimodal.py:66).
"""
import os

import torch
from torch import nn

from . import ops
from .layers import BidirectionalLSTM, DenseSoftmax
from .losses import DROP_FRAMES, KERAS_CTC_EPS, _prep_lengths, _raise_on_status


class UnimodalNet(nn.Module):
    def __init__(self, numfeats, units, nb_classes, noise_std, dropouts, seed=47):
        super().__init__()
        self.numfeats, self.units, self.nb_classes, self.noise_std = numfeats, units, nb_classes, noise_std
        self.p1, self.p2, self.pd = dropouts
        self.blstm_1 = BidirectionalLSTM(numfeats, units, dropout=self.p1, seed=seed, name="blstm_1")
        self.blstm_2 = BidirectionalLSTM(2 * units, units, dropout=self.p2, seed=seed + 7, name="blstm_2")
        self.dense = DenseSoftmax(2 * units, nb_classes, seed=seed)

    @property
    def layers(self):
        """Keras depth-sorted layer list as multimodal.py:109-118 indexes it."""
        return [None, None, self.blstm_1, self.blstm_2]

    def sample_regularisers(self, B, T, seed, step, device):
        reg = {}
        off = int(step) * 64
        if self.noise_std > 0:
            reg["noise"] = ops.gaussian_noise((B, T, self.numfeats), self.noise_std, seed, off, device)
        if self.p1 > 0:
            reg["m1"] = ops.dropout_mask((8, B, self.numfeats), self.p1, seed + 1, off, device)
        if self.p2 > 0:
            reg["m2"] = ops.dropout_mask((8, B, 2 * self.units), self.p2, seed + 2, off, device)
        if self.pd > 0:
            reg["drop"] = ops.dropout_mask((B, T, 2 * self.units), self.pd, seed + 3, off, device)
        return reg

    def tower(self, x, reg=None, merged=None, col0=0):
        """GaussianNoise -> BLSTM -> BLSTM -> residual add.  With `merged` (a (B,T,Fo) buffer, inference
        only) the add lands in columns [col0, col0 + 2*units) of it -- the fusion model's concat."""
        reg = reg or {}
        if reg.get("noise") is not None:
            x = ops.add(x, reg["noise"])
        y1 = self.blstm_1(x, reg.get("m1"))
        y2 = self.blstm_2(y1, reg.get("m2"))
        if merged is not None:
            return ops.add_into(y1, y2, merged, col0)
        return _add(y1, y2)

    def forward(self, x, reg=None):
        """-> (y_pred softmax probabilities (B,T,C), inner logits)."""
        reg = reg or {}
        res = self.tower(x, reg)
        return self.dense(res, reg.get("drop"))


class _AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return ops.add(a, b)

    @staticmethod
    def backward(ctx, g):
        return g, g


def _add(a, b):
    return _AddFn.apply(a, b)


def SpeechNet(numfeats=39, nb_classes=44, units=500, seed=47):
    return UnimodalNet(numfeats, units, nb_classes, noise_std=0.5, dropouts=(0.4, 0.5, 0.5), seed=seed)


def SkeletalNet(numfeats=20, nb_classes=22, units=300, seed=53):
    return UnimodalNet(numfeats, units, nb_classes, noise_std=0.5, dropouts=(0.6, 0.6, 0.6), seed=seed)


class FusionNet(nn.Module):
    def __init__(self, speech=None, skeletal=None, nb_classes=22, units=100, seed=61):
        super().__init__()
        self.speech = speech if speech is not None else SpeechNet()
        self.skeletal = skeletal if skeletal is not None else SkeletalNet()
        # multimodal.py:135-148: freeze the four tower Bidirectional layers
        for tower in (self.speech, self.skeletal):
            for l in (tower.blstm_1, tower.blstm_2):
                l.trainable = True
                l.forward_layer.trainable = False
                l.backward_layer.trainable = False
            for p in tower.dense.parameters():  # the uni-modal heads are not part of the fusion graph
                p.requires_grad_(False)
        fin = 2 * self.speech.units + 2 * self.skeletal.units
        self.units, self.nb_classes = units, nb_classes
        self._streams = None
        self.blstm_3 = BidirectionalLSTM(fin, units, dropout=0.5, seed=seed, name="blstm_2")
        self.dense = DenseSoftmax(2 * units, nb_classes, seed=seed)

    def trainable_parameters(self):
        return [self.blstm_3.kernel, self.blstm_3.recurrent_kernel, self.blstm_3.bias, self.dense.kernel,
                self.dense.bias]

    def sample_regularisers(self, B, T, seed, step, device):
        off = int(step) * 64
        reg = {"sp": {}, "sk": {}}
        sp = self.speech.sample_regularisers(B, T, seed + 10, step, device)
        sk = self.skeletal.sample_regularisers(B, T, seed + 20, step, device)
        sp.pop("drop", None)
        sk.pop("drop", None)
        sk.pop("noise", None)  # GaussianNoise(0.0) on the skeletal branch (multimodal.py:105)
        reg["sp"], reg["sk"] = sp, sk
        reg["m3"] = ops.dropout_mask((8, B, self.blstm_3.input_dim), 0.5, seed + 30, off, device)
        reg["drop"] = ops.dropout_mask((B, T, 2 * self.units), 0.5, seed + 31, off, device)
        return reg

    def merged(self, xa, xs, reg=None):
        reg = reg or {}
        with torch.no_grad():
            # the two towers are independent until the concat (multimodal.py:109-118,155): run them on
            # two streams -- their recurrences (64 + 40 persistent CTAs) and GEMMs overlap on the 148 SMs
            fa, fs = 2 * self.speech.units, 2 * self.skeletal.units
            fused = fa % 4 == 0 and fs % 4 == 0   # 16-byte column blocks: the adds write the concat directly
            merged = torch.empty(xa.shape[:2] + (fa + fs,), dtype=torch.float32, device=xa.device) if fused else None
            if os.environ.get("GR_TOWER_STREAMS", "1") == "0":
                ra = self.speech.tower(xa, reg.get("sp"), merged, 0)
                rs = self.skeletal.tower(xs, reg.get("sk"), merged, fa)
                return merged if fused else ops.concat2(ra, rs)
            cur = torch.cuda.current_stream()
            if self._streams is None:
                self._streams = (torch.cuda.Stream(), torch.cuda.Stream())
            sa, sb = self._streams
            sa.wait_stream(cur)
            sb.wait_stream(cur)
            with torch.cuda.stream(sa):
                ra = self.speech.tower(xa, reg.get("sp"), merged, 0)
            with torch.cuda.stream(sb):
                rs = self.skeletal.tower(xs, reg.get("sk"), merged, fa)
            cur.wait_stream(sa)
            cur.wait_stream(sb)
            return merged if fused else ops.concat2(ra, rs)

    def forward(self, xa, xs, reg=None):
        reg = reg or {}
        m = self.merged(xa, xs, reg)
        y3 = self.blstm_3(m, reg.get("m3"))
        return self.dense(y3, reg.get("drop"))

    # ------------------------------------------------------------------ explicit training step
    def loss_and_grads(self, xa, xs, labels, input_length, label_length, reg=None, global_batch=None,
                       check=False, eps=KERAS_CTC_EPS):
        """One forward+backward of the Keras objective mean_b ctc_loss_b (dummy loss, speech:131).
        Returns (loss (B,), [grads in trainable_parameters() order]).  `global_batch` = divisor of
        the mean (the data-parallel global batch; default = local B)."""
        reg = reg or {}
        B, T = xa.shape[0], xa.shape[1]
        gb = float(global_batch or B)
        lab_i, il, ll = _prep_lengths(labels, input_length, label_length, xa.device)
        m = self.merged(xa, xs, reg)
        l3 = self.blstm_3
        y3 = l3(m, reg.get("m3"))                      # autograd node (kernel-backed)
        H2 = 2 * self.units
        y3_2d = y3.detach().reshape(B * T, H2)
        dm = reg.get("drop")
        dm2 = None if dm is None else dm.reshape(B * T, H2)
        logits, _ = ops.dense_softmax_fwd(y3_2d, self.dense.kernel.detach(), self.dense.bias.detach(), dm2,
                                          want_probs=False)
        upstream = torch.full((B,), 1.0 / gb, dtype=torch.float32, device=xa.device)
        loss, g_logits, status = ops.ctc_loss_grad(logits.reshape(B, T, self.nb_classes), lab_i, ll, il, True,
                                                   drop_frames=DROP_FRAMES, eps=eps, upstream=upstream)
        if check:
            _raise_on_status(status)
        dWd, dbd, dy3 = ops.dense_bwd(y3_2d, self.dense.kernel.detach(), g_logits.reshape(B * T, -1), dm2,
                                      want_dx=True)
        gW, gU, gb_ = torch.autograd.grad(y3, [l3.kernel, l3.recurrent_kernel, l3.bias],
                                          grad_outputs=dy3.reshape(B, T, H2))
        return loss, [gW, gU, gb_, dWd, dbd]


class KerasAdam:
    """Adam(lr, clipvalue, decay) + maxnorm(3) on LSTM `kernel`s, as compiled at multimodal.py:206-213."""

    def __init__(self, params, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-7, decay=0.0, clipvalue=0.0,
                 maxnorm_params=(), max_norm=3.0):
        self.params = list(params)
        self.lr, self.beta1, self.beta2, self.eps, self.decay, self.clipvalue = lr, beta1, beta2, eps, decay, clipvalue
        self.maxnorm_ids = {id(p) for p in maxnorm_params}
        self.max_norm = max_norm
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.iterations = 0

    def step(self, grads):
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            mn = self.max_norm if id(p) in self.maxnorm_ids else 0.0
            ops.adam_step(p.data, g, m, v, self.iterations, self.lr, self.beta1, self.beta2, self.eps, self.decay,
                          self.clipvalue, mn)
        self.iterations += 1


def fusion_optimizer(model):
    """Adam(lr=0.0001, clipvalue=0.5, decay=1e-5) over the fusion BLSTM + Dense (multimodal.py:206-208)."""
    return KerasAdam(model.trainable_parameters(), lr=1e-4, clipvalue=0.5, decay=1e-5,
                     maxnorm_params=[model.blstm_3.kernel], max_norm=3.0)
