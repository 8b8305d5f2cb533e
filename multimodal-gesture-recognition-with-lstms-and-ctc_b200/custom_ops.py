"""`torch.library` registration of the hot path: `torch.ops.mgr_b200.*`.

BASELINE.json's north_star words the host side as "PyTorch custom ops calling hand-written sm_100a CUDA through a thin
C-ABI extension".  The modules in this package (`losses`, `layers`, `sequence_decoding`) call the C ABI from
`autograd.Function`s; this file registers the SAME kernels as dispatcher ops, each with a fake (meta) kernel and, where
the reference differentiates through it, an autograd formula, so that they are visible to fake-tensor tracing,
`torch.library.opcheck` and export:

  mgr_b200::ctc_loss(x, labels, label_length, input_length, input_is_logits, drop_frames, eps)
        -> (loss (B,1), grad_x (B,T,C), status (B,))          K.ctc_batch_cost / losses.py:4-15, differentiable wrt x
  mgr_b200::blstm_forward(x, W, U, b, masks?, mask_scale, keep) -> (y (B,T,2H), gates, cell)
  mgr_b200::blstm_backward(...) -> (dx, dW, dU, db)     Bidirectional(LSTM) (speech_lstm_ctc_words.py:56-77): forward is
        differentiable wrt x, W, U, b (autograd formula = blstm_backward); `custom_ops.blstm` picks `keep`
  mgr_b200::ctc_bestpath(probs, threshold, drop_frames) -> (ids, lens)           sequence_decoding.py:38-69
  mgr_b200::ctc_greedy(probs, seq_len?, eps) -> (ids, lens, neg_sum_logits)       K.ctc_decode(greedy=True)
  mgr_b200::ctc_beam(probs, seq_len?, beam_width, top_paths, eps) -> (ids, lens, log_prob)   K.ctc_decode(greedy=False)

CUDA only (`device_types="cuda"`): there is no CPU kernel, a CPU tensor raises from the dispatcher.
"""
from typing import Optional, Tuple

import torch
from torch import Tensor
from torch.library import custom_op

from . import ops


# ----------------------------------------------------------------------------------- CTC loss
@custom_op("mgr_b200::ctc_loss", mutates_args=(), device_types="cuda")
def ctc_loss(x: Tensor, labels: Tensor, label_length: Tensor, input_length: Tensor, input_is_logits: bool,
             drop_frames: int, eps: float) -> Tuple[Tensor, Tensor, Tensor]:
    loss, grad, status = ops.ctc_loss_grad(x, labels.to(torch.int32).contiguous(),
                                           label_length.reshape(-1).to(torch.int32).contiguous(),
                                           input_length.reshape(-1).to(torch.int32).contiguous(), input_is_logits,
                                           drop_frames=drop_frames, eps=eps)
    return loss.reshape(-1, 1), grad, status


@ctc_loss.register_fake
def _(x, labels, label_length, input_length, input_is_logits, drop_frames, eps):
    B = x.shape[0]
    return x.new_empty((B, 1)), torch.empty_like(x), x.new_empty((B,), dtype=torch.int32)


def _ctc_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1])


def _ctc_backward(ctx, g_loss, g_grad, g_status):
    (grad,) = ctx.saved_tensors
    gx = grad * g_loss.reshape(-1, 1, 1) if g_loss is not None else None
    return gx, None, None, None, None, None, None


ctc_loss.register_autograd(_ctc_backward, setup_context=_ctc_setup)


# ----------------------------------------------------------------------------------- BLSTM
@custom_op("mgr_b200::blstm_forward", mutates_args=(), device_types="cuda")
def blstm_forward(x: Tensor, W: Tensor, U: Tensor, b: Tensor, masks: Optional[Tensor], mask_scale: float,
                  keep: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """-> (y, activated gates (B*T, 8H), cell (B,T,2H)); the last two are empty when keep is False."""
    from .layers import _project
    B, T, F = x.shape
    H = U.shape[1]
    x, W, U, b = x.contiguous(), W.contiguous(), U.contiguous(), b.contiguous()
    masks = None if masks is None else masks.contiguous()
    gates = _project(x.reshape(B * T, F), W, b, masks, B, T, H, mask_scale=mask_scale)
    y, cell = ops.lstm_recurrence_fwd(gates, U, B, T, H, keep_cell=keep)
    if not keep:
        return y, x.new_empty((0,)), x.new_empty((0,))
    return y, gates, cell


@blstm_forward.register_fake
def _(x, W, U, b, masks, mask_scale, keep):
    B, T, _ = x.shape
    H = U.shape[1]
    y = x.new_empty((B, T, 2 * H))
    if not keep:
        return y, x.new_empty((0,)), x.new_empty((0,))
    return y, x.new_empty((B * T, 8 * H)), x.new_empty((B, T, 2 * H))


@custom_op("mgr_b200::blstm_backward", mutates_args=(), device_types="cuda")
def blstm_backward(x: Tensor, W: Tensor, U: Tensor, masks: Optional[Tensor], gates: Tensor, cell: Tensor, y: Tensor,
                   dy: Tensor, mask_scale: float, need_dx: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """BPTT + the weight-gradient contractions -> (dx or empty, dW (F,8H), dU (2,H,4H), db (8H))."""
    B, T, F = x.shape
    H = U.shape[1]
    BT = B * T
    dP = ops.lstm_recurrence_bwd(gates.clone(), cell, dy.contiguous(), U.contiguous(), B, T, H).reshape(BT, 8 * H)
    x2, y2 = x.contiguous().reshape(BT, F), y.reshape(BT, 2 * H)
    db = ops.colsum(dP)
    dpt_hi, dpt_lo = ops.split_bf16(dP, transpose=True)
    dW = torch.empty((F, 8 * H), dtype=torch.float32, device=x.device)
    if masks is None:
        ops.gemm_a32(x2, dpt_hi, dpt_lo, F, 8 * H, BT, dW, 8 * H, transA=True, rows_per_seq=T)
    else:
        ops.gemm_a32(x2, dpt_hi, dpt_lo, F, H, BT, dW, 8 * H, nvar=8, mask=masks.contiguous(), rows_per_seq=T, transA=True,
                     mask_scale=mask_scale)
    dU = torch.empty((2, H, 4 * H), dtype=torch.float32, device=x.device)
    for d in range(2):
        ops.gemm_a32(y2, dpt_hi[d * 4 * H:(d + 1) * 4 * H], dpt_lo[d * 4 * H:(d + 1) * 4 * H], H, 4 * H, BT, dU[d], 4 * H,
                     transA=True, rows_per_seq=T, row_shift=-1 if d == 0 else 1, a_col_offset=d * H)
    if not need_dx:
        return x.new_empty((0,)), dW, dU, db
    dx2 = torch.empty((BT, F), dtype=torch.float32, device=x.device)
    Wc = W.contiguous()
    if masks is None:
        w_hi, w_lo = ops.split_bf16(Wc)
        ops.gemm_a32(dP, w_hi, w_lo, BT, F, 8 * H, dx2, F)
    else:
        tmp = torch.empty((BT, F), dtype=torch.float32, device=x.device)
        for dg in range(8):
            w_hi, w_lo = ops.split_bf16(Wc, ncols=H, col_offset=dg * H)
            ops.gemm_a32(dP, w_hi, w_lo, BT, F, H, tmp, F, a_col_offset=dg * H)
            ops.mask_mul_acc(dx2, tmp, masks[dg].contiguous(), T, accumulate=dg > 0)
    return dx2.reshape(B, T, F), dW, dU, db


@blstm_backward.register_fake
def _(x, W, U, masks, gates, cell, y, dy, mask_scale, need_dx):
    H = U.shape[1]
    F = x.shape[2]
    return (torch.empty_like(x) if need_dx else x.new_empty((0,)), x.new_empty((F, 8 * H)), x.new_empty((2, H, 4 * H)),
            x.new_empty((8 * H,)))


def _blstm_setup(ctx, inputs, output):
    x, W, U, b, masks, mask_scale, keep = inputs
    y, gates, cell = output
    ctx.keep, ctx.has_masks, ctx.mask_scale = keep, masks is not None, mask_scale
    ctx.save_for_backward(x, W, U, gates, cell, y, *([] if masks is None else [masks]))


def _blstm_backward(ctx, dy, _g_gates, _g_cell):
    if not ctx.keep:
        raise RuntimeError("mgr_b200::blstm_forward was called with keep=False: nothing was saved for backward")
    x, W, U, gates, cell, y, *m = ctx.saved_tensors
    dx, dW, dU, db = torch.ops.mgr_b200.blstm_backward(x, W, U, m[0] if ctx.has_masks else None, gates, cell, y,
                                                       dy.contiguous(), ctx.mask_scale, ctx.needs_input_grad[0])
    return (dx if ctx.needs_input_grad[0] else None), dW, dU, db, None, None, None


blstm_forward.register_autograd(_blstm_backward, setup_context=_blstm_setup)


def blstm(x, W, U, b, masks=None, mask_scale=0.0):
    """Bidirectional(LSTM) through the dispatcher ops: keeps the activations only when a gradient can be asked for."""
    keep = torch.is_grad_enabled() and any(t.requires_grad for t in (x, W, U, b))
    return torch.ops.mgr_b200.blstm_forward(x, W, U, b, masks, float(mask_scale), keep)[0]


# ----------------------------------------------------------------------------------- decoders
@custom_op("mgr_b200::ctc_bestpath", mutates_args=(), device_types="cuda")
def ctc_bestpath(probs: Tensor, threshold: float, drop_frames: int) -> Tuple[Tensor, Tensor]:
    return ops.bestpath_ref(probs, threshold, drop_frames)


@ctc_bestpath.register_fake
def _(probs, threshold, drop_frames):
    N, T, _ = probs.shape
    return probs.new_empty((N, T), dtype=torch.int32), probs.new_empty((N,), dtype=torch.int32)


@custom_op("mgr_b200::ctc_greedy", mutates_args=(), device_types="cuda")
def ctc_greedy(probs: Tensor, seq_len: Optional[Tensor], eps: float) -> Tuple[Tensor, Tensor, Tensor]:
    sl = None if seq_len is None else seq_len.reshape(-1).to(torch.int32).contiguous()
    return ops.greedy(probs, sl, eps)


@ctc_greedy.register_fake
def _(probs, seq_len, eps):
    N, T, _ = probs.shape
    return (probs.new_empty((N, T), dtype=torch.int32), probs.new_empty((N,), dtype=torch.int32), probs.new_empty((N,)))


@custom_op("mgr_b200::ctc_beam", mutates_args=(), device_types="cuda")
def ctc_beam(probs: Tensor, seq_len: Optional[Tensor], beam_width: int, top_paths: int,
             eps: float) -> Tuple[Tensor, Tensor, Tensor]:
    sl = None if seq_len is None else seq_len.reshape(-1).to(torch.int32).contiguous()
    return ops.beam(probs, sl, beam_width=beam_width, top_paths=top_paths, eps=eps)


@ctc_beam.register_fake
def _(probs, seq_len, beam_width, top_paths, eps):
    N, T, _ = probs.shape
    return (probs.new_empty((N, top_paths, T), dtype=torch.int32), probs.new_empty((N, top_paths), dtype=torch.int32),
            probs.new_empty((N, top_paths)))
