"""`ctc_lambda_func` -- the reference's loss contract on the B200 CTC kernel.

Mirrors /root/reference/audio_network/losses.py:4-15 (== multimodal_fusion/losses.py:4-15):
    y_pred, labels, input_length, label_length = args
    y_pred = y_pred[:, 2:, :]
    return K.ctc_batch_cost(labels, y_pred, input_length, label_length)      # (B, 1)
Same argument order and meaning (y_pred = softmax probabilities (B,T,C); labels float (B,Lmax)
padded with -1; input_length (B,1) counted AFTER the two dropped frames; label_length (B,1));
same error behaviour (the conditions TF reports as InvalidArgumentError raise
`InvalidArgumentError`); differentiable wrt y_pred.  `softmax_ctc` is the fused form that takes
the Dense layer's logits and returns the gradient wrt the logits directly.
"""
import torch

from . import ops
from ._lib import CTC_STATUS, InvalidArgumentError

KERAS_CTC_EPS = 1e-8  # literal in the pinned Keras 2.1.4 (later Keras: K.epsilon() = 1e-7)
DROP_FRAMES = 2       # losses.py:11


def _prep_lengths(labels, input_length, label_length, device):
    lab = torch.as_tensor(labels)
    B = lab.shape[0]
    lab_i = lab.to(device=device).to(torch.int32)  # float -> int32 truncation, as K.ctc_label_dense_to_sparse
    il = torch.as_tensor(input_length).reshape(B).to(device=device, dtype=torch.int32)
    ll = torch.as_tensor(label_length).reshape(B).to(device=device, dtype=torch.int32)
    return lab_i.contiguous(), il.contiguous(), ll.contiguous()


def _raise_on_status(status):
    st = status.cpu()
    bad = ((st != 0) & (st != 4)).nonzero()
    if bad.numel():
        b = int(bad[0])
        raise InvalidArgumentError(CTC_STATUS.get(int(st[b]), "CTC error in batch %d") % b)


class _CtcFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, lab_i, ll, il, is_logits, drop, eps, check):
        loss, grad, status = ops.ctc_loss_grad(x, lab_i, ll, il, is_logits, drop_frames=drop, eps=eps,
                                               want_grad=x.requires_grad or True)
        if check:
            _raise_on_status(status)
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(status)
        return loss.reshape(-1, 1), status

    @staticmethod
    def backward(ctx, g_loss, _g_status):
        (grad,) = ctx.saved_tensors
        # d(sum_b g_b * loss_b)/dx ; broadcasting multiply is plumbing, not a kernel of the path
        return grad * g_loss.reshape(-1, 1, 1), None, None, None, None, None, None, None


def ctc_batch_cost(y_true, y_pred, input_length, label_length, eps=KERAS_CTC_EPS, check=True):
    """Keras `K.ctc_batch_cost(y_true, y_pred, input_length, label_length)` -> (B, 1)."""
    lab_i, il, ll = _prep_lengths(y_true, input_length, label_length, y_pred.device)
    loss, _ = _CtcFn.apply(y_pred, lab_i, ll, il, False, 0, eps, check)
    return loss


def ctc_lambda_func(args, eps=KERAS_CTC_EPS, check=True):
    """Drop-in for the reference's `ctc_lambda_func(args)`; returns the (B, 1) loss tensor."""
    y_pred, labels, input_length, label_length = args
    lab_i, il, ll = _prep_lengths(labels, input_length, label_length, y_pred.device)
    # the [:, 2:, :] slice is applied inside the kernel (drop_frames) -- no copy
    loss, _ = _CtcFn.apply(y_pred, lab_i, ll, il, False, DROP_FRAMES, eps, check)
    return loss


def softmax_ctc(inner_logits, labels, input_length, label_length, eps=KERAS_CTC_EPS, check=True):
    """Fused Activation('softmax') + ctc_lambda_func on the Dense output (speech:86-109)."""
    lab_i, il, ll = _prep_lengths(labels, input_length, label_length, inner_logits.device)
    loss, _ = _CtcFn.apply(inner_logits, lab_i, ll, il, True, DROP_FRAMES, eps, check)
    return loss
