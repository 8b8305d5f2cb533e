"""Oracle: Keras-2.1.4 ``Bidirectional(LSTM(...), merge_mode='concat')`` and the three
reference topologies, restated with torch CPU tensors (autograd supplies the backward).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Parity unpinned by reference fixtures;
pinned by fp64 finite differences (autograd gradcheck) and hand-computed single steps.

Call sites followed:
  * /root/reference/audio_network/speech_lstm_ctc_words.py:32-134 (39 -> BLSTM500 x2 -> add ->
    Dropout -> Dense44 -> softmax -> CTC),
  * /root/reference/skeletal_network/skeletal_lstm_ctc.py:296-394 (20 -> BLSTM300 x2 -> add ->
    Dropout -> Dense22 -> softmax -> CTC),
  * /root/reference/multimodal_fusion/multimodal.py:58-215 (frozen towers -> concat(speech,
    skeletal) -> BLSTM100 -> Dropout -> Dense22 -> softmax -> CTC; freeze logic :33-55,:135-148).
Keras LSTMCell maths: gate order i,f,c,o on the 4H axis,
recurrent_activation hard_sigmoid(v)=clip(0.2v+0.5,0,1), activation tanh, h0=c0=0, no masking.
"""
import numpy as np
import torch


def hard_sigmoid(v):
    return torch.clamp(0.2 * v + 0.5, 0.0, 1.0)


def keras_lstm(x, kernel, recurrent, bias, go_backwards=False, masks=None):
    """One Keras LSTM(return_sequences=True).  x (B,T,F); kernel (F,4H); recurrent (H,4H);
    bias (4H).  masks: None or (4,B,F) input-dropout masks (already scaled 1/(1-p)), one per
    gate i,f,c,o, constant over time.  For go_backwards the output is in *processing* order
    (Keras semantics); Bidirectional re-reverses it."""
    B, T, F = x.shape
    H = recurrent.shape[0]
    if go_backwards:
        x = torch.flip(x, dims=[1])
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    outs = []
    for t in range(T):
        xt = x[:, t, :]
        if masks is None:
            zx = xt @ kernel + bias
        else:
            zx = torch.cat([(xt * masks[g]) @ kernel[:, g * H:(g + 1) * H] for g in range(4)], dim=1) + bias
        z = zx + h @ recurrent
        i = hard_sigmoid(z[:, 0 * H:1 * H])
        f = hard_sigmoid(z[:, 1 * H:2 * H])
        g = torch.tanh(z[:, 2 * H:3 * H])
        o = hard_sigmoid(z[:, 3 * H:4 * H])
        c = f * c + i * g
        h = o * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, dim=1)


def bidirectional_lstm(x, weights, masks_fwd=None, masks_bwd=None):
    """Bidirectional(LSTM, merge_mode='concat').  weights in Keras get_weights() order:
    [fwd kernel, fwd recurrent, fwd bias, bwd kernel, bwd recurrent, bwd bias]."""
    fk, fr, fb, bk, br, bb = weights
    yf = keras_lstm(x, fk, fr, fb, go_backwards=False, masks=masks_fwd)
    yb = keras_lstm(x, bk, br, bb, go_backwards=True, masks=masks_bwd)
    yb = torch.flip(yb, dims=[1])
    return torch.cat([yf, yb], dim=2)


def dense_softmax(x, w, b):
    a = x @ w + b
    return torch.softmax(a, dim=-1), a


# ------------------------------------------------------------------ weight initialisers
def orthogonal(rng, rows, cols):
    a = rng.standard_normal((max(rows, cols), min(rows, cols)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    q = q if rows >= cols else q.T
    return q[:rows, :cols]


def init_blstm_weights(rng, F, H, dtype=np.float32):
    """kernel ~ U(-0.05,0.05) (RandomUniform, speech:42-44,63), recurrent orthogonal, bias zeros
    with the forget slice at one (unit_forget_bias)."""
    w = []
    for _ in range(2):
        k = rng.uniform(-0.05, 0.05, size=(F, 4 * H))
        r = orthogonal(rng, H, 4 * H)
        b = np.zeros(4 * H)
        b[H:2 * H] = 1.0
        w += [k.astype(dtype), r.astype(dtype), b.astype(dtype)]
    return w


def init_dense_weights(rng, F, C, dtype=np.float32):
    return [rng.uniform(-0.05, 0.05, size=(F, C)).astype(dtype), np.zeros(C, dtype=dtype)]


# ------------------------------------------------------------------ topologies
def unimodal_forward(x, w1, w2, wd, noise=None, masks=None, drop_mask=None):
    """speech_lstm_ctc_words.py:48-90 / skeletal_lstm_ctc.py:300-350.
    noise: additive GaussianNoise sample (B,T,F) or None.  masks: dict with optional keys
    'l1f','l1b','l2f','l2b' -> (4,B,F) input-dropout masks.  drop_mask: (B,T,2H) Dropout mask
    (scaled) or None.  Returns (softmax probs, logits, residual sum)."""
    masks = masks or {}
    if noise is not None:
        x = x + noise
    y1 = bidirectional_lstm(x, w1, masks.get("l1f"), masks.get("l1b"))
    y2 = bidirectional_lstm(y1, w2, masks.get("l2f"), masks.get("l2b"))
    res = y1 + y2
    d = res if drop_mask is None else res * drop_mask
    p, a = dense_softmax(d, wd[0], wd[1])
    return p, a, res


def tower_forward(x, w1, w2, noise=None, masks=None):
    """The re-used tower of multimodal.py:109-118: BLSTM1 -> BLSTM2 -> add (no head)."""
    masks = masks or {}
    if noise is not None:
        x = x + noise
    y1 = bidirectional_lstm(x, w1, masks.get("l1f"), masks.get("l1b"))
    y2 = bidirectional_lstm(y1, w2, masks.get("l2f"), masks.get("l2b"))
    return y1 + y2


def fusion_forward(xa, xs, sp_w1, sp_w2, sk_w1, sk_w2, fu_w, fu_d, noise_a=None, masks=None,
                   drop_mask=None):
    """multimodal.py:103-179.  Speech features first in the concat (:155-156).  Towers carry no
    gradient (frozen, :135-148): computed under no_grad."""
    masks = masks or {}
    with torch.no_grad():
        ra = tower_forward(xa, sp_w1, sp_w2, noise=noise_a,
                           masks={k[3:]: v for k, v in masks.items() if k.startswith("sp_")})
        rs = tower_forward(xs, sk_w1, sk_w2,
                           masks={k[3:]: v for k, v in masks.items() if k.startswith("sk_")})
        merged = torch.cat([ra, rs], dim=2)
    y3 = bidirectional_lstm(merged, fu_w, masks.get("fu_f"), masks.get("fu_b"))
    d = y3 if drop_mask is None else y3 * drop_mask
    p, a = dense_softmax(d, fu_d[0], fu_d[1])
    return p, a, merged


def torch_ctc_lambda(y_pred, labels, input_length, label_length, eps=1e-8):
    """Differentiable torch restatement of losses.py:4-15 used for whole-model gradient oracles
    and as the CPU baseline: F.ctc_loss(log_softmax(log(p[:,2:]+eps))) with blank=C-1 and the
    'label >= C-1 terminates' rule applied first.  Independent of oracle/ctc_ref.py (which is
    checked against it)."""
    import torch.nn.functional as Fn
    B, T, C = y_pred.shape
    z = torch.log(y_pred[:, 2:, :] + eps)
    lsm = torch.log_softmax(z, dim=-1).transpose(0, 1)  # (T', B, C)
    il = torch.as_tensor(input_length).reshape(B).to(torch.long)
    ll = torch.as_tensor(label_length).reshape(B).to(torch.long)
    lab = torch.as_tensor(labels)
    tgt = []
    tl = []
    for b in range(B):
        seq = []
        for v in lab[b, :int(ll[b])].tolist():
            if int(v) >= C - 1:
                break
            seq.append(int(v))
        tgt += seq
        tl.append(len(seq))
    tgt = torch.tensor(tgt, dtype=torch.long)
    tl = torch.tensor(tl, dtype=torch.long)
    loss = Fn.ctc_loss(lsm, tgt, il, tl, blank=C - 1, reduction="none", zero_infinity=False)
    return loss.reshape(B, 1)
