"""Keras-surface layers on the B200 kernels.

`BidirectionalLSTM(F, H, dropout=p)` mirrors
    Bidirectional(LSTM(H, activation='tanh', recurrent_activation='hard_sigmoid',
                       recurrent_dropout=0.0, dropout=p, kernel_constraint=maxnorm(3),
                       kernel_initializer=RandomUniform(-0.05, 0.05, seed=47),
                       return_sequences=True), merge_mode='concat')
(/root/reference/audio_network/speech_lstm_ctc_words.py:56-77, skeletal_network/
skeletal_lstm_ctc.py:309-331, multimodal_fusion/multimodal.py:159-168): (B,T,F) -> (B,T,2H),
`get_weights()/set_weights()` in Keras order [fwd kernel (F,4H), fwd recurrent (H,4H), fwd bias
(4H), bwd kernel, bwd recurrent, bwd bias], gate order i,f,c,o, `.trainable` for freezing
(multimodal.py:33-55).  `DenseSoftmax` mirrors Dense(C)+Activation('softmax') (speech:86-90).
"""
import os

import numpy as np
import torch
from torch import nn

from . import ops

GEMM_PASSES = 3  # bf16x3 (fp32-faithful) tensor-core projections; 1 = plain bf16
BIAS_COLUMN = os.environ.get("GR_BIAS_COLUMN", "1") != "0"   # see _project


def _project(x2, W, b, masks, B, T, H, passes=3, mask_scale=0.0):
    """Hoisted input projection P = (X o mask_g) W_g + b -> (B*T, 8H) on tcgen05.  X stays fp32 in
    HBM: masking and the bf16 hi/lo split are fused into the GEMM's producer warps.
    `mask_scale` > 0: the masks are dropout masks with values in {0, mask_scale}."""
    BT, F = x2.shape
    gates = torch.empty((BT, 8 * H), dtype=torch.float32, device=x2.device)
    if b is not None and BIAS_COLUMN and H > 128 and F + 1 <= 64:
        # Narrow first layers (K = 39 / 20: ONE k-block, a pure store stream of B*T x 8H floats): the bias rides in the
        # k-block's padding -- a column of ones appended to x, the bias as the matching row of W -- so the store-bound
        # epilogue has nothing to add (1.0 * (b_hi + b_lo) reproduces b to 2^-17 relative).  H > 128 keeps the call on
        # the one-variant-per-CTA schedule, where the mask is a plain multiply (its ones-column stays 1).
        Fp = (F + 1 + 3) // 4 * 4
        xp = torch.nn.functional.pad(x2, (0, Fp - F))
        xp[:, F] = 1.0
        Wp = torch.zeros((Fp, 8 * H), dtype=torch.float32, device=W.device)
        Wp[:F] = W
        Wp[F] = b
        mp = None if masks is None else torch.nn.functional.pad(masks, (0, Fp - F), value=1.0)
        wt_hi, wt_lo = ops.split_bf16(Wp, transpose=True)
        if mp is None:
            ops.gemm_a32(xp, wt_hi, wt_lo, BT, 8 * H, Fp, gates, 8 * H)
        else:
            ops.gemm_a32(xp, wt_hi, wt_lo, BT, H, Fp, gates, 8 * H, nvar=8, mask=mp, rows_per_seq=T)
        return gates
    wt_hi, wt_lo = ops.split_bf16(W, transpose=True)  # (8H, pad8(F)): B operand, K-major
    if F % 4:
        # 39 MFCC / odd feature counts: zero-pad the (small) input to 16-byte rows so that the
        # vectorised producer path applies; the padded weight columns are zero as well
        Fp = (F + 3) // 4 * 4
        x2 = torch.nn.functional.pad(x2, (0, Fp - F))
        if masks is not None:
            masks = torch.nn.functional.pad(masks, (0, Fp - F))
        F = Fp
    if masks is None:
        ops.gemm_a32(x2, wt_hi, wt_lo, BT, 8 * H, F, gates, 8 * H, bias=b)
    else:
        ops.gemm_a32(x2, wt_hi, wt_lo, BT, H, F, gates, 8 * H, nvar=8, mask=masks, rows_per_seq=T, bias=b,
                     mask_scale=mask_scale)
    return gates


_SIDE = {}   # device -> two side streams for half-batch recurrences


def _halves(B, H, device):
    """Training recurrences of a layer hold grid(B,H) SMs (38 at H=300, 64 at H=500) for T dependent steps while the rest
    of the GPU idles -- nothing else of a uni-modal training step can run beside them.  When two half-batch launches fit
    side by side they run on two streams: each half's step is shorter (smaller tiles) and both proceed concurrently
    (config-2 layer, fwd + BPTT: 8.33 -> 7.05 ms).  Results are bit-identical (rows are independent).  None = do not split."""
    if os.environ.get("GR_TRAIN_SPLIT", "1") == "0" or B < 64 or B % 2:
        return None
    g = ops.lstm_recurrence_grid(B // 2, H)
    if g <= 0 or 2 * g > torch.cuda.get_device_properties(device).multi_processor_count:
        return None
    if device not in _SIDE:
        _SIDE[device] = (torch.cuda.Stream(device=device), torch.cuda.Stream(device=device))
    return _SIDE[device]


def _recurrence_fwd_train(gates, U, B, T, H):
    """Forward recurrence keeping gates / cell for BPTT -> (y, cell); two concurrent half batches when they fit."""
    side = _halves(B, H, gates.device)
    if side is None:
        return ops.lstm_recurrence_fwd(gates, U, B, T, H, keep_cell=True)
    y = torch.empty((B, T, 2 * H), dtype=torch.float32, device=gates.device)
    cell = torch.empty((B, T, 2 * H), dtype=torch.float32, device=gates.device)
    cur, h = torch.cuda.current_stream(), B // 2
    for i, s_ in enumerate(side):
        s_.wait_stream(cur)
        with torch.cuda.stream(s_):
            ops.lstm_recurrence_fwd(gates[i * h * T:(i + 1) * h * T], U, h, T, H, keep_cell=True, y=y[i * h:(i + 1) * h],
                                    cell=cell[i * h:(i + 1) * h])
    for s_ in side:
        cur.wait_stream(s_)
    return y, cell


def _recurrence_bwd(gates, cell, dy, U, B, T, H):
    """BPTT in place over `gates` -> dP (B*T, 8H); two concurrent half batches when they fit."""
    side = _halves(B, H, gates.device)
    dy = dy.contiguous()
    if side is None:
        return ops.lstm_recurrence_bwd(gates, cell, dy, U, B, T, H).reshape(B * T, 8 * H)
    cur, h = torch.cuda.current_stream(), B // 2
    g2 = gates.reshape(B * T, 8 * H)
    for i, s_ in enumerate(side):
        s_.wait_stream(cur)
        with torch.cuda.stream(s_):
            ops.lstm_recurrence_bwd(g2[i * h * T:(i + 1) * h * T], cell[i * h:(i + 1) * h], dy[i * h:(i + 1) * h], U, h, T, H)
    for s_ in side:
        cur.wait_stream(s_)
    return g2


class _BlstmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, U, b, masks, passes, mask_scale=0.0):
        B, T, F = x.shape
        H = U.shape[1]
        # raw pointers cross the C ABI: every operand must be dense row-major
        x, W, U, b = x.contiguous(), W.contiguous(), U.contiguous(), b.contiguous()
        masks = None if masks is None else masks.contiguous()
        x2 = x.reshape(B * T, F)
        need_grad = any(ctx.needs_input_grad[:4])
        gates = _project(x2, W, b, masks, B, T, H, mask_scale=mask_scale)
        if need_grad:
            y, cell = _recurrence_fwd_train(gates, U, B, T, H)
        else:
            y, cell = ops.lstm_recurrence_fwd(gates, U, B, T, H, keep_cell=False)
        ctx.mask_scale = mask_scale
        if need_grad:
            ctx.save_for_backward(x, W, U, masks, gates, cell, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, U, masks, gates, cell, y = ctx.saved_tensors
        B, T, F = x.shape
        H = U.shape[1]
        BT = B * T
        dP = _recurrence_bwd(gates, cell, dy, U, B, T, H)
        x2 = x.reshape(BT, F)
        y2 = y.reshape(BT, 2 * H)
        db = ops.colsum(dP) if ctx.needs_input_grad[3] else None
        dW = dU = dx = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dpt_hi, dpt_lo = ops.split_bf16(dP, transpose=True)  # dP^T (8H, pad8(BT)): B operand of dW and dU
        if ctx.needs_input_grad[1]:
            # dW_g = (X o mask_g)^T dP_g : A = X read transposed by the producer warps
            dW = torch.empty((F, 8 * H), dtype=torch.float32, device=x.device)
            if masks is None:
                ops.gemm_a32(x2, dpt_hi, dpt_lo, F, 8 * H, BT, dW, 8 * H, transA=True, rows_per_seq=T)
            else:
                ops.gemm_a32(x2, dpt_hi, dpt_lo, F, H, BT, dW, 8 * H, nvar=8, mask=masks, rows_per_seq=T, transA=True,
                             mask_scale=ctx.mask_scale)
        if ctx.needs_input_grad[2]:
            # dU_d = H_prev^T dP_d : A = y shifted by one step inside each sequence
            dU = torch.empty((2, H, 4 * H), dtype=torch.float32, device=x.device)
            for d in range(2):
                ops.gemm_a32(y2, dpt_hi[d * 4 * H:(d + 1) * 4 * H], dpt_lo[d * 4 * H:(d + 1) * 4 * H], H, 4 * H, BT,
                             dU[d], 4 * H, transA=True, rows_per_seq=T, row_shift=-1 if d == 0 else 1,
                             a_col_offset=d * H)
        if ctx.needs_input_grad[0]:
            dx2 = torch.empty((BT, F), dtype=torch.float32, device=x.device)
            if masks is None:
                w_hi, w_lo = ops.split_bf16(W)  # (F, 8H) K-major
                ops.gemm_a32(dP, w_hi, w_lo, BT, F, 8 * H, dx2, F)
            else:
                tmp = torch.empty((BT, F), dtype=torch.float32, device=x.device)
                for dg in range(8):
                    n0 = dg * H
                    w_hi, w_lo = ops.split_bf16(W, ncols=H, col_offset=n0)
                    ops.gemm_a32(dP, w_hi, w_lo, BT, F, H, tmp, F, a_col_offset=n0)
                    ops.mask_mul_acc(dx2, tmp, masks[dg], T, accumulate=dg > 0)
            dx = dx2.reshape(B, T, F)
        return dx, dW, dU, db, None, None, None


def blstm_into(x, W, U, b, masks, mask_scale, aux, col0, accumulate, want_y):
    """Inference-only BLSTM whose output ALSO lands in columns [col0, col0 + 2H) of `aux` (B,T,Fo): stored, or added to
    what is there (`accumulate`).  Returns y, or None with `want_y=False`.  Used by the frozen towers of the fusion
    model: layer 1 stores into the Merge(concat) buffer, layer 2 accumulates -- the residual add costs no extra pass."""
    B, T, F = x.shape
    H = U.shape[1]
    with torch.no_grad():
        x, W, U, b = x.contiguous(), W.contiguous(), U.contiguous(), b.contiguous()
        masks = None if masks is None else masks.contiguous()
        gates = _project(x.reshape(B * T, F), W, b, masks, B, T, H, mask_scale=mask_scale)
        y, _ = ops.lstm_recurrence_fwd(gates, U, B, T, H, keep_cell=False, aux=aux, aux_col0=col0,
                                       aux_accumulate=accumulate, want_y=want_y)
    return y


def blstm(x, W, U, b, masks=None, passes=None, mask_scale=0.0):
    """Functional form.  W (F,8H) = [fwd kernel | bwd kernel]; U (2,H,4H); b (8H);
    masks None or (8, B, F): input-dropout masks for (dir, gate) = (0,i),(0,f),(0,c),(0,o),(1,i)..;
    mask_scale > 0 declares them Keras dropout masks, every element 0 or mask_scale = 1/(1-rate)
    (lets the projection kernels share one bf16 split between the four gates of a direction)."""
    return _BlstmFn.apply(x, W, U, b, masks, GEMM_PASSES if passes is None else passes, float(mask_scale))


class _DirView:
    """`.forward_layer` / `.backward_layer` handle carrying the `.trainable` flag (multimodal.py:47-48)."""

    def __init__(self, owner):
        self._owner = owner
        self._trainable = True

    @property
    def trainable(self):
        return self._trainable

    @trainable.setter
    def trainable(self, v):
        self._trainable = bool(v)
        self._owner._sync_trainable()


class BidirectionalLSTM(nn.Module):
    def __init__(self, input_dim, units, dropout=0.0, max_norm=3.0, seed=47, name=None):
        super().__init__()
        self.input_dim, self.units, self.dropout, self.max_norm = input_dim, units, float(dropout), max_norm
        self.name = name
        F, H = input_dim, units
        rng = np.random.default_rng(seed)
        W = rng.uniform(-0.05, 0.05, size=(F, 8 * H)).astype(np.float32)
        U = np.stack([_orthogonal(rng, H, 4 * H) for _ in range(2)]).astype(np.float32)
        b = np.zeros(8 * H, dtype=np.float32)
        b[H:2 * H] = 1.0          # unit_forget_bias, forward layer
        b[5 * H:6 * H] = 1.0      # backward layer
        self.kernel = nn.Parameter(torch.from_numpy(np.ascontiguousarray(W)))
        self.recurrent_kernel = nn.Parameter(torch.from_numpy(np.ascontiguousarray(U)))
        self.bias = nn.Parameter(torch.from_numpy(np.ascontiguousarray(b)))
        self.trainable = True
        self.forward_layer = _DirView(self)
        self.backward_layer = _DirView(self)
        self.training_phase = True  # K.set_learning_phase(1): dropout active (speech:40)

    def _sync_trainable(self):
        # the reference freezes by flipping forward_layer/backward_layer.trainable (multimodal.py:43-48)
        on = self.forward_layer.trainable and self.backward_layer.trainable
        for p in (self.kernel, self.recurrent_kernel, self.bias):
            p.requires_grad_(on)

    def get_weights(self):
        H = self.units
        W, U, b = self.kernel.detach().cpu().numpy(), self.recurrent_kernel.detach().cpu().numpy(), \
            self.bias.detach().cpu().numpy()
        return [W[:, :4 * H].copy(), U[0].copy(), b[:4 * H].copy(), W[:, 4 * H:].copy(), U[1].copy(),
                b[4 * H:].copy()]

    def set_weights(self, weights):
        fk, fr, fb, bk, br, bb = [np.asarray(w, dtype=np.float32) for w in weights]
        dev = self.kernel.device
        with torch.no_grad():
            self.kernel.copy_(torch.from_numpy(np.concatenate([fk, bk], axis=1)).to(dev))
            self.recurrent_kernel.copy_(torch.from_numpy(np.stack([fr, br])).to(dev))
            self.bias.copy_(torch.from_numpy(np.concatenate([fb, bb])).to(dev))

    def make_masks(self, B, seed, offset):
        """4 masks per direction, shape (B,F), scaled 1/(1-p), constant over time."""
        if self.dropout <= 0.0:
            return None
        return ops.dropout_mask((8, B, self.input_dim), self.dropout, seed, offset, self.kernel.device)

    def _scale(self, masks, dropout_masks):
        return 1.0 / (1.0 - self.dropout) if (dropout_masks and masks is not None and 0.0 < self.dropout < 1.0) else 0.0

    def forward_into(self, x, masks, dropout_masks, aux, col0, accumulate, want_y=True):
        """No-grad forward that also stores / accumulates the output into `aux[..., col0:col0+2*units]` (see blstm_into)."""
        return blstm_into(x, self.kernel, self.recurrent_kernel, self.bias, masks, self._scale(masks, dropout_masks), aux,
                          col0, accumulate, want_y)

    def forward(self, x, masks=None, dropout_masks=False):
        """`dropout_masks=True`: `masks` came from `make_masks` / `ops.dropout_mask` with this layer's rate, i.e. every
        element is 0 or 1/(1-rate) (injected masks with other values must leave it False)."""
        return blstm(x, self.kernel, self.recurrent_kernel, self.bias, masks, mask_scale=self._scale(masks, dropout_masks))


class _DenseSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, Wd, bd, drop_mask):
        shp = x.shape
        x2 = x.contiguous().reshape(-1, shp[-1])
        m2 = None if drop_mask is None else drop_mask.contiguous().reshape(-1, shp[-1])
        Wd, bd = Wd.contiguous(), bd.contiguous()
        logits, probs = ops.dense_softmax_fwd(x2, Wd, bd, m2)
        ctx.save_for_backward(x2, Wd, m2, probs)
        ctx.shp = shp
        C = Wd.shape[1]
        return probs.reshape(shp[:-1] + (C,)), logits.reshape(shp[:-1] + (C,))

    @staticmethod
    def backward(ctx, g_probs, g_logits):
        x2, Wd, m2, probs = ctx.saved_tensors
        C = Wd.shape[1]
        g = torch.zeros_like(probs)
        if g_probs is not None:
            gp = g_probs.reshape(-1, C)
            g = g + probs * (gp - (probs * gp).sum(dim=1, keepdim=True))  # softmax backward (unfused path)
        if g_logits is not None:
            g = g + g_logits.reshape(-1, C)
        dW, db, dx = ops.dense_bwd(x2, Wd, g.contiguous(), m2, want_dx=ctx.needs_input_grad[0])
        return (None if dx is None else dx.reshape(ctx.shp)), dW, db, None


class DenseSoftmax(nn.Module):
    """Dense(nb_classes, kernel_initializer=RandomUniform(+-0.05, seed=47)) + Activation('softmax')."""

    def __init__(self, input_dim, nb_classes, seed=47):
        super().__init__()
        rng = np.random.default_rng(seed + 1)
        self.kernel = nn.Parameter(torch.from_numpy(rng.uniform(-0.05, 0.05, size=(input_dim, nb_classes)).astype(np.float32)))
        self.bias = nn.Parameter(torch.zeros(nb_classes))

    def get_weights(self):
        return [self.kernel.detach().cpu().numpy().copy(), self.bias.detach().cpu().numpy().copy()]

    def set_weights(self, weights):
        with torch.no_grad():
            self.kernel.copy_(torch.as_tensor(np.asarray(weights[0], dtype=np.float32)).to(self.kernel.device))
            self.bias.copy_(torch.as_tensor(np.asarray(weights[1], dtype=np.float32)).to(self.bias.device))

    def forward(self, x, drop_mask=None):
        """Returns (softmax probabilities, logits)."""
        return _DenseSoftmaxFn.apply(x, self.kernel, self.bias, drop_mask)


def _orthogonal(rng, rows, cols):
    a = rng.standard_normal((max(rows, cols), min(rows, cols)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    q = q if rows >= cols else q.T
    return q[:rows, :cols]
