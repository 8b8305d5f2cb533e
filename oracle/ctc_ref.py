"""Oracle: CTC loss + gradient with the semantics of the reference's loss op.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Parity unpinned by reference
fixtures; pinned by brute-force enumeration / finite differences / torch CPU.

Follows, in order:
  * /root/reference/audio_network/losses.py:4-15 (== multimodal_fusion/losses.py:4-15;
    copies at skeletal_network/skeletal_lstm_ctc.py:257-268): ``ctc_lambda_func`` --
    unpack (y_pred, labels, input_length, label_length), drop frames 0 and 1,
    call ``K.ctc_batch_cost``.
  * Keras 2.1.4 ``K.ctc_batch_cost`` (requirements.txt:4; not vendored): labels[:label_length],
    ``log(transpose(y_pred) + eps)``, ``tf.nn.ctc_loss`` with defaults, expand_dims(.,1).
  * TensorFlow 1.12.1 ``CTCLossCalculator`` (requirements.txt:8; not vendored): softmax of
    the op input, blank = C-1, alpha/beta in log space, loss = -log p(l|x),
    grad = y - exp(logsumexp_{u:l'[u]=c}(alpha+beta) - log p).  
"""
import itertools

import numpy as np

KERAS_CTC_EPS = 1e-8  # literal in Keras 2.1.x ctc_batch_cost (later Keras: K.epsilon()=1e-7)


class CTCInvalidArgument(ValueError):
    """Mirrors tf.errors.InvalidArgumentError raised by CTCLossOp at run time."""


def prepare_label_sequence(raw_labels, num_classes):
    """TF CTCLossCalculator label rule: a label >= C-1 ends the sequence; a
    non-null label after a null one is an error.  `raw_labels` = labels[b, :label_length[b]]."""
    out = []
    finished = False
    for v in raw_labels:
        v = int(v)  # float -> int32 by truncation (K.ctc_label_dense_to_sparse + to_int32)
        if v >= num_classes - 1:
            finished = True
        elif finished:
            raise CTCInvalidArgument(
                "Saw a non-null label (index >= num_classes - 1) following a null label")
        else:
            if v < 0:
                raise CTCInvalidArgument("label out of range")
            out.append(v)
    return out


def _lse(a, b):
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.logaddexp(a, b)


def ctc_loss_grad_single(z, label_seq, seq_len, dtype=np.float64):
    """tf.nn.ctc_loss for ONE batch element.

    z: (T, C) op inputs (pre-softmax activations); label_seq: prepared list (may be empty);
    seq_len: frames used.  Returns (loss, dz (T, C), alpha (U, Tn), beta (U, Tn)).
    dz rows >= seq_len are zero.  No valid path -> loss=+inf, dz = softmax(z) (TF behaviour).
    """
    z = np.asarray(z, dtype=dtype)
    T, C = z.shape
    Tn = int(seq_len)
    blank = C - 1
    lp = [blank]
    for l in label_seq:
        lp += [int(l), blank]
    lp = np.asarray(lp, dtype=np.int64)
    U = lp.shape[0]
    zz = z[:Tn]
    m = zz.max(axis=1, keepdims=True)
    e = np.exp(zz - m)
    y = (e / e.sum(axis=1, keepdims=True)).astype(dtype)  # linear softmax, as TF
    with np.errstate(divide="ignore"):
        logy = np.log(y).astype(dtype)
    ninf = dtype(-np.inf)
    alpha = np.full((U, Tn), ninf, dtype=dtype)
    beta = np.full((U, Tn), ninf, dtype=dtype)
    # skip-transition mask: l'[u] != blank and l'[u] != l'[u-2]
    skip = np.zeros(U, dtype=bool)
    skip[2:] = (lp[2:] != blank) & (lp[2:] != lp[:-2])
    skip_b = np.zeros(U, dtype=bool)  # for beta: l'[u] != blank and l'[u] != l'[u+2]
    skip_b[:-2] = (lp[:-2] != blank) & (lp[:-2] != lp[2:])

    alpha[0, 0] = logy[0, blank]
    if U > 1:
        alpha[1, 0] = logy[0, lp[1]]
    for t in range(1, Tn):
        prev = alpha[:, t - 1]
        s = prev.copy()
        s[1:] = _lse(s[1:], prev[:-1])
        s2 = np.full(U, ninf, dtype=dtype)
        s2[2:] = prev[:-2]
        s = np.where(skip, _lse(s, s2), s)
        lo = max(0, U - 2 * (Tn - t))
        hi = min(U, 2 * (t + 1))
        a = logy[t, lp] + s
        a[:lo] = ninf
        a[hi:] = ninf
        alpha[:, t] = a.astype(dtype)
    beta[U - 1, Tn - 1] = 0
    if U > 1:
        beta[U - 2, Tn - 1] = 0
    for t in range(Tn - 2, -1, -1):
        nxt = beta[:, t + 1] + logy[t + 1, lp]
        s = nxt.copy()
        s[:-1] = _lse(s[:-1], nxt[1:])
        s2 = np.full(U, ninf, dtype=dtype)
        s2[:-2] = nxt[2:]
        s = np.where(skip_b, _lse(s, s2), s)
        lo = max(0, U - 2 * (Tn - t))
        hi = min(U, 2 * (t + 1))
        s[:lo] = ninf
        s[hi:] = ninf
        beta[:, t] = s.astype(dtype)
    ab0 = alpha[:, 0] + beta[:, 0]
    logp = ninf
    for u in range(U):
        logp = _lse(logp, ab0[u])
    dz = np.zeros((T, C), dtype=dtype)
    if not np.isfinite(logp):
        dz[:Tn] = y
        return dtype(np.inf), dz, alpha, beta
    ab = alpha + beta  # (U, Tn)
    occ = np.full((Tn, C), ninf, dtype=dtype)
    for u in range(U):
        occ[:, lp[u]] = _lse(occ[:, lp[u]], ab[u])
    with np.errstate(under="ignore"):
        dz[:Tn] = y - np.exp(occ - logp)
    return dtype(-logp), dz, alpha, beta


def ctc_batch_cost(y_true, y_pred, input_length, label_length, eps=KERAS_CTC_EPS,
                   dtype=np.float64, want_grad=True):
    """Keras ``K.ctc_batch_cost`` restated.  y_pred (B, T', C) probabilities.

    Returns loss (B, 1) and, if want_grad, d(loss_b)/d(y_pred[b]) (B, T', C).
    Raises CTCInvalidArgument like the TF op (label_length 0, label_length > input_length, ...).
    """
    y_pred = np.asarray(y_pred, dtype=dtype)
    y_true = np.asarray(y_true)
    B, T, C = y_pred.shape
    il = np.asarray(input_length).reshape(B).astype(np.int64)
    ll = np.asarray(label_length).reshape(B).astype(np.int64)
    loss = np.zeros((B, 1), dtype=dtype)
    grad = np.zeros_like(y_pred) if want_grad else None
    for b in range(B):
        if il[b] > T:
            raise CTCInvalidArgument("sequence_length(%d) <= %d" % (b, T))
        if ll[b] <= 0:
            raise CTCInvalidArgument("Labels length is zero in batch %d" % b)
        raw = y_true[b, :ll[b]]
        seq = prepare_label_sequence(raw, C)
        if ll[b] > il[b]:
            raise CTCInvalidArgument(
                "Not enough time for target transition sequence (required: %d, available: %d)"
                % (ll[b], il[b]))
        z = np.log(y_pred[b] + dtype(eps)).astype(dtype)
        l, dz, _, _ = ctc_loss_grad_single(z, seq, il[b], dtype=dtype)
        loss[b, 0] = l
        if want_grad:
            grad[b] = dz / (y_pred[b] + dtype(eps))
    return (loss, grad) if want_grad else loss


def ctc_lambda_func(args, eps=KERAS_CTC_EPS, dtype=np.float64, want_grad=False):
    """/root/reference/audio_network/losses.py:4-15 restated on NumPy arrays.

    args = (y_pred (B,T,C) softmax probabilities, labels (B,Lmax) float, input_length (B,1),
    label_length (B,1)).  Frames 0 and 1 are dropped (losses.py:11); input_length counts frames
    AFTER the drop (data_generator.py:223 feeds T-2).  Returns (B,1) loss [and grad wrt the
    full y_pred, zero on frames 0,1]."""
    y_pred, labels, input_length, label_length = args
    y_pred = np.asarray(y_pred, dtype=dtype)
    out = ctc_batch_cost(labels, y_pred[:, 2:, :], input_length, label_length, eps=eps,
                         dtype=dtype, want_grad=want_grad)
    if not want_grad:
        return out
    loss, g = out
    full = np.zeros_like(y_pred)
    full[:, 2:, :] = g
    return loss, full


def softmax_ctc_grad_logits(logits, labels, input_length, label_length, eps=KERAS_CTC_EPS,
                            dtype=np.float64, upstream=None):
    """Gradient chain of SURVEY.md A.3: model head softmax -> ctc_lambda_func.

    Returns (loss (B,1), g_a (B,T,C)) where g_a = d(sum_b upstream_b * loss_b)/d(logits);
    upstream defaults to 1/B (the Keras objective = batch mean of the Lambda output,
    speech_lstm_ctc_words.py:131)."""
    a = np.asarray(logits, dtype=dtype)
    B = a.shape[0]
    m = a.max(axis=2, keepdims=True)
    e = np.exp(a - m)
    p = e / e.sum(axis=2, keepdims=True)
    loss, g_p = ctc_lambda_func((p, labels, input_length, label_length), eps=eps, dtype=dtype,
                                want_grad=True)
    up = np.full(B, 1.0 / B, dtype=dtype) if upstream is None else np.asarray(upstream, dtype).reshape(B)
    g_p = g_p * up[:, None, None]
    g_a = p * (g_p - (p * g_p).sum(axis=2, keepdims=True))
    return loss, g_a


# ----------------------------------------------------------------------------- pins
def collapse_path(path, blank):
    out = []
    prev = None
    for c in path:
        if c != prev and c != blank:
            out.append(c)
        prev = c
    return out


def brute_force_neg_log_prob(y, label_seq):
    """-log sum over all C^T alignments that collapse to label_seq.  y: (T, C) probabilities."""
    y = np.asarray(y, dtype=np.float64)
    T, C = y.shape
    blank = C - 1
    tot = 0.0
    target = list(label_seq)
    for path in itertools.product(range(C), repeat=T):
        if collapse_path(path, blank) == target:
            pr = 1.0
            for t, c in enumerate(path):
                pr *= y[t, c]
            tot += pr
    return -np.log(tot) if tot > 0 else np.inf
