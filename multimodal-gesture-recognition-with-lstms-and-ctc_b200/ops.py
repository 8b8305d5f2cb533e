"""Thin functional wrappers: torch CUDA tensors in, C-ABI calls (include/gr_b200.h) out.

torch is used for device memory and streams only; every computation below is one of the
hand-written sm_100a kernels in csrc/.  No CPU fallback: non-CUDA tensors raise.
"""
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr, require_cuda

_ws_cache = {}


def _workspace(key, nbytes, device):
    """Grow-only per-(purpose, device, stream) scratch buffer (torch caching allocator memory)."""
    k = (key, device, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(k)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _ws_cache[k] = buf
    return buf


def _pad8(n):
    return (n + 7) // 8 * 8


def _f32c(t):
    if t.dtype != torch.float32:
        raise _lib.GrError("expected float32 tensor, got %s" % t.dtype)
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------------- CTC
def ctc_loss_grad(x, labels_i32, label_len_i32, input_len_i32, input_is_logits, drop_frames=2,
                  eps=1e-8, upstream=None, want_grad=True):
    """Returns (loss (B,), grad (B,T,C) or None, status (B,) int32)."""
    require_cuda(x, labels_i32, label_len_i32, input_len_i32, upstream)
    x = _f32c(x)
    B, T, C = x.shape
    Lmax = labels_i32.shape[1]
    import ctypes
    nbytes = ctypes.c_size_t(0)
    call("gr_ctc_workspace_bytes", B, T, C, Lmax, ctypes.byref(nbytes))
    ws = _workspace("ctc", nbytes.value, x.device)
    loss = torch.empty(B, dtype=torch.float32, device=x.device)
    grad = torch.empty_like(x) if want_grad else None
    status = torch.empty(B, dtype=torch.int32, device=x.device)
    call("gr_ctc_loss_grad_f32", ptr(x), int(bool(input_is_logits)), B, T, C, int(drop_frames), float(eps),
         ptr(labels_i32), Lmax, ptr(label_len_i32), ptr(input_len_i32), ptr(upstream), ptr(loss), ptr(grad),
         ptr(status), ptr(ws), ws.numel(), stream_ptr())
    return loss, grad, status


# ----------------------------------------------------------------------------------- decode
def bestpath_ref(probs, threshold, drop_frames=2):
    require_cuda(probs)
    probs = _f32c(probs)
    N, T, C = probs.shape
    ids = torch.empty((N, T), dtype=torch.int32, device=probs.device)
    lens = torch.empty(N, dtype=torch.int32, device=probs.device)
    call("gr_ctc_bestpath_ref_f32", ptr(probs), N, T, C, int(drop_frames), float(threshold), ptr(ids), ptr(lens),
         stream_ptr())
    return ids, lens


def greedy(probs, seq_len=None, eps=1e-8):
    require_cuda(probs, seq_len)
    probs = _f32c(probs)
    N, T, C = probs.shape
    ids = torch.empty((N, T), dtype=torch.int32, device=probs.device)
    lens = torch.empty(N, dtype=torch.int32, device=probs.device)
    score = torch.empty(N, dtype=torch.float32, device=probs.device)
    call("gr_ctc_greedy_f32", ptr(probs), N, T, C, ptr(seq_len), float(eps), ptr(ids), ptr(lens), ptr(score),
         stream_ptr())
    return ids, lens, score


def beam(probs, seq_len=None, beam_width=100, top_paths=1, merge_repeated=True, eps=1e-8):
    import ctypes
    require_cuda(probs, seq_len)
    probs = _f32c(probs)
    N, T, C = probs.shape
    nbytes = ctypes.c_size_t(0)
    call("gr_ctc_beam_workspace_bytes", N, T, C, int(beam_width), ctypes.byref(nbytes))
    ws = _workspace("beam", nbytes.value, probs.device)
    ids = torch.empty((N, top_paths, T), dtype=torch.int32, device=probs.device)
    lens = torch.empty((N, top_paths), dtype=torch.int32, device=probs.device)
    logp = torch.empty((N, top_paths), dtype=torch.float32, device=probs.device)
    call("gr_ctc_beam_f32", ptr(probs), N, T, C, ptr(seq_len), float(eps), int(beam_width), int(top_paths),
         int(bool(merge_repeated)), ptr(ids), ptr(lens), ptr(logp), ptr(ws), ws.numel(), stream_ptr())
    return ids, lens, logp


# ----------------------------------------------------------------------------------- GEMM
def split_bf16(x2d, mask=None, rows_per_seq=1, add=None, noise=None, transpose=False, row_shift=0,
               ncols=None, col_offset=0, want_lo=True):
    """fp32 (R, ldx) -> (hi, lo) bf16.  Uses columns [col_offset, col_offset+ncols) of x2d.
    not transpose: out (R, pad8(ncols)); transpose: out (ncols, pad8(R))."""
    require_cuda(x2d, mask, add, noise)
    assert x2d.dim() == 2 and x2d.stride(1) == 1
    R, ldx = x2d.shape[0], x2d.stride(0)
    K = x2d.shape[1] - col_offset if ncols is None else ncols
    shape = (K, _pad8(R)) if transpose else (R, _pad8(K))
    hi = torch.empty(shape, dtype=torch.bfloat16, device=x2d.device)
    lo = torch.empty(shape, dtype=torch.bfloat16, device=x2d.device) if want_lo else None
    base = x2d.data_ptr() + 4 * col_offset
    import ctypes
    off = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr() + 4 * col_offset)
    call("gr_split_bf16_f32", ctypes.c_void_p(base), off(add), off(noise), ptr(mask), int(rows_per_seq), R, K,
         ldx, int(bool(transpose)), int(row_shift), ptr(hi), ptr(lo), shape[1], stream_ptr())
    return hi, lo


def gemm_nt(a_hi, a_lo, b_hi, b_lo, M, N, K, out=None, ldc=None, bias=None, accumulate=False, passes=3,
            out_col_offset=0):
    """C[M,N] (+)= A[M,K] B[N,K]^T (+bias) on tcgen05; operands pre-split (hi, lo) bf16, K-major.
    K is the padded width (multiple of 8).  out may be a wider matrix (ldc, column offset)."""
    import ctypes
    require_cuda(a_hi, b_hi, out, bias)
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a_hi.device)
        ldc = N
    cptr = ctypes.c_void_p(out.data_ptr() + 4 * out_col_offset)
    _lib.note_work(2.0 * M * N * K)
    call("gr_gemm_bf16x3_f32", ptr(a_hi), ptr(a_lo), ptr(b_hi), ptr(b_lo), ptr(bias), cptr, int(ldc), int(M),
         int(N), int(K), a_hi.stride(0), b_hi.stride(0), int(passes), int(bool(accumulate)), stream_ptr())
    return out


def gemm_a32(A2d, b_hi, b_lo, M, Nv, K, out, ldc, nvar=1, mask=None, rows_per_seq=1, transA=False, row_shift=0,
             bias=None, accumulate=False, a_col_offset=0, out_col_offset=0, mask_scale=0.0):
    """Fused-prologue projection GEMM (gr_gemm_a32_f32): A stays fp32 in HBM; mask / hi-lo split /
    transpose / time shift happen in the kernel's producer warps.  A2d: (rows, lda) fp32.
    `mask_scale` > 0 declares `mask` a dropout mask with values in {0, mask_scale} (gr_gemm_a32_dropout_f32)."""
    import ctypes
    require_cuda(A2d, b_hi, b_lo, out, bias, mask)
    assert A2d.dim() == 2 and A2d.stride(1) == 1
    aptr = ctypes.c_void_p(A2d.data_ptr() + 4 * a_col_offset)
    cptr = ctypes.c_void_p(out.data_ptr() + 4 * out_col_offset)
    _lib.note_work(2.0 * M * Nv * nvar * K)
    if mask is not None and mask_scale > 0.0:
        call("gr_gemm_a32_dropout_f32", aptr, A2d.stride(0), int(bool(transA)), int(row_shift), ptr(mask),
             float(mask_scale), int(rows_per_seq), int(nvar), ptr(b_hi), ptr(b_lo), b_hi.stride(0), ptr(bias), cptr,
             int(ldc), int(M), int(Nv), int(K), int(bool(accumulate)), stream_ptr())
        return out
    call("gr_gemm_a32_f32", aptr, A2d.stride(0), int(bool(transA)), int(row_shift), ptr(mask), int(rows_per_seq),
         int(nvar), ptr(b_hi), ptr(b_lo), b_hi.stride(0), ptr(bias), cptr, int(ldc), int(M), int(Nv), int(K),
         int(bool(accumulate)), stream_ptr())
    return out


def gemm_simt(A, B, bias=None, out=None, accumulate=False):
    require_cuda(A, B, bias, out)
    M, K = A.shape
    N = B.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    call("gr_gemm_simt_f32", ptr(A), ptr(B), ptr(bias), ptr(out), M, N, K, A.stride(0), B.stride(0),
         out.stride(0), int(bool(accumulate)), stream_ptr())
    return out


# ----------------------------------------------------------------------------------- LSTM
def lstm_workspace(B, H, device):
    import ctypes
    nbytes = ctypes.c_size_t(0)
    call("gr_lstm_workspace_bytes", int(B), int(H), ctypes.byref(nbytes))
    return _workspace("lstm", nbytes.value, device)


def lstm_recurrence_grid(B, H):
    """CTAs of the tensor-memory recurrence for this shape (0: another kernel runs)."""
    return int(_lib.load().gr_lstm_recurrence_grid(int(B), int(H)))


def lstm_aux_supported(B, H):
    """True when the recurrence of this shape can write / accumulate h into an auxiliary (B,T,ld) buffer."""
    return bool(_lib.load().gr_lstm_recurrence_aux_supported(int(B), int(H)))


def lstm_recurrence_fwd(gates, U, B, T, H, keep_cell=True, aux=None, aux_col0=0, aux_accumulate=False, want_y=True,
                        y=None, cell=None):
    """`aux`: (B, T, Fo) buffer whose columns [aux_col0, aux_col0 + 2H) also receive h (stored, or added when
    `aux_accumulate`); with `want_y=False` no separate y tensor is written (returns y = None).  `y` / `cell`:
    caller-provided contiguous (B,T,2H) outputs (e.g. the two halves of one tensor)."""
    import ctypes
    require_cuda(gates, U, aux, y, cell)
    if y is None:
        y = torch.empty((B, T, 2 * H), dtype=torch.float32, device=gates.device) if (want_y or aux is None) else None
    if cell is None:
        cell = torch.empty((B, T, 2 * H), dtype=torch.float32, device=gates.device) if keep_cell else None
    ws = lstm_workspace(B, H, gates.device)
    # algorithmic bytes: read 4H pre-activations, write h (+ 4H gates + c when kept)
    _lib.note_work(float(B) * T * 2 * H * (40 if keep_cell else 20))
    if aux is not None:
        assert aux.dim() == 3 and aux.shape[0] == B and aux.shape[1] == T and aux.is_contiguous()
        aptr = ctypes.c_void_p(aux.data_ptr() + 4 * int(aux_col0))
        call("gr_lstm_recurrence_fwd_aux_f32", ptr(gates), ptr(U), B, T, H, ptr(y), ptr(cell), aptr, int(aux.shape[2]),
             int(bool(aux_accumulate)), ptr(ws), ws.numel(), stream_ptr())
        return y, cell
    call("gr_lstm_recurrence_fwd_f32", ptr(gates), ptr(U), B, T, H, ptr(y), ptr(cell), ptr(ws), ws.numel(),
         stream_ptr())
    return y, cell


def lstm_recurrence_bwd(gates, cell, dy, U, B, T, H):
    require_cuda(gates, cell, dy, U)
    ws = lstm_workspace(B, H, gates.device)
    _lib.note_work(float(B) * T * 2 * H * 44)  # read gates 16 + c,c_prev 8 + dy 4, write dP 16
    call("gr_lstm_recurrence_bwd_f32", ptr(gates), ptr(cell), ptr(_f32c(dy)), ptr(U), B, T, H, ptr(ws),
         ws.numel(), stream_ptr())
    return gates  # now dP


# ----------------------------------------------------------------------------------- head etc.
def dense_softmax_fwd(x2d, Wd, bd, drop_mask=None, want_logits=True, want_probs=True):
    require_cuda(x2d, Wd, bd, drop_mask)
    R, Fin = x2d.shape
    C = Wd.shape[1]
    logits = torch.empty((R, C), dtype=torch.float32, device=x2d.device) if want_logits else None
    probs = torch.empty((R, C), dtype=torch.float32, device=x2d.device) if want_probs else None
    call("gr_dense_softmax_fwd_f32", ptr(x2d), ptr(drop_mask), ptr(Wd), ptr(bd), R, Fin, C, ptr(logits),
         ptr(probs), stream_ptr())
    return logits, probs


def dense_bwd(x2d, Wd, g_logits, drop_mask=None, want_dx=False):
    require_cuda(x2d, Wd, g_logits, drop_mask)
    R, Fin = x2d.shape
    C = Wd.shape[1]
    dW = torch.empty((Fin, C), dtype=torch.float32, device=x2d.device)
    db = torch.empty(C, dtype=torch.float32, device=x2d.device)
    dx = torch.empty((R, Fin), dtype=torch.float32, device=x2d.device) if want_dx else None
    call("gr_dense_bwd_f32", ptr(x2d), ptr(drop_mask), ptr(Wd), ptr(_f32c(g_logits)), R, Fin, C, ptr(dW), ptr(db),
         ptr(dx), stream_ptr())
    return dW, db, dx


def colsum(a2d, out=None):
    require_cuda(a2d)
    R, N = a2d.shape
    if out is None:
        out = torch.empty(N, dtype=torch.float32, device=a2d.device)
    call("gr_colsum_f32", ptr(a2d), R, N, a2d.stride(0), ptr(out), stream_ptr())
    return out


def add(a, b, out=None):
    require_cuda(a, b)
    a, b = _f32c(a), _f32c(b)
    if out is None:
        out = torch.empty_like(a)
    call("gr_add_f32", ptr(a), ptr(b), ptr(out), a.numel(), stream_ptr())
    return out


def add_into(a, b, out, col0):
    """out[..., col0:col0+F] = a + b with `out` a wider (.., Fo) buffer: the tower's residual add written into
    its column block of the Merge(concat) buffer (gr_add_into_f32)."""
    import ctypes
    require_cuda(a, b, out)
    a, b = _f32c(a), _f32c(b)
    F, Fo = a.shape[-1], out.shape[-1]
    if not out.is_contiguous():
        raise _lib.GrError("add_into: out must be contiguous")
    rows = a.numel() // F
    call("gr_add_into_f32", ptr(a), ptr(b), ctypes.c_void_p(out.data_ptr() + 4 * col0), rows, F, Fo, stream_ptr())
    return out


def concat2(a, b):
    require_cuda(a, b)
    a, b = _f32c(a), _f32c(b)
    Fa, Fb = a.shape[-1], b.shape[-1]
    rows = a.numel() // Fa
    out = torch.empty(a.shape[:-1] + (Fa + Fb,), dtype=torch.float32, device=a.device)
    call("gr_concat2_f32", ptr(a), Fa, ptr(b), Fb, ptr(out), rows, stream_ptr())
    return out


def mask_mul_acc(out, tmp, mask, rows_per_seq, accumulate):
    R, K = tmp.shape
    call("gr_mask_mul_acc_f32", ptr(out), ptr(tmp), ptr(mask), int(rows_per_seq), R, K, int(bool(accumulate)),
         stream_ptr())
    return out


def adam_step(param, grad, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-7, decay=0.0, clipvalue=0.0,
              max_norm=0.0):
    require_cuda(param, grad, m, v)
    rows, cols = (param.shape[0], param.numel() // param.shape[0]) if param.dim() >= 2 else (1, param.numel())
    call("gr_adam_step_f32", ptr(param), ptr(_f32c(grad)), ptr(m), ptr(v), param.numel(), rows, cols, float(lr),
         float(beta1), float(beta2), float(eps), float(decay), float(clipvalue), float(max_norm), int(step),
         stream_ptr())


def _ptr_table(tensors):
    """Host arrays (device pointers, element counts) for the multi-tensor entry points."""
    import ctypes
    n = len(tensors)
    ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in tensors])
    sizes = (ctypes.c_size_t * n)(*[t.numel() for t in tensors])
    return ptrs, sizes, n


def pack(tensors, flat):
    """flat <- the tensors back to back, one launch (the data-parallel gradient bucket)."""
    require_cuda(flat, *tensors)
    ts = [_f32c(t) for t in tensors]
    ptrs, sizes, n = _ptr_table(ts)
    call("gr_pack_f32", ptrs, sizes, n, ptr(flat), stream_ptr())
    return flat


def adam_flat(params, flat_grad, flat_m, flat_v, step, lr, beta1=0.9, beta2=0.999, eps=1e-7, decay=0.0, clipvalue=0.0):
    """clipvalue + Keras Adam on every tensor of `params` in ONE launch; gradients / moments are flat buffers holding the
    tensors back to back in the same order."""
    require_cuda(flat_grad, flat_m, flat_v, *params)
    ptrs, sizes, n = _ptr_table(params)
    call("gr_adam_flat_f32", ptrs, sizes, n, ptr(flat_grad), ptr(flat_m), ptr(flat_v), float(lr), float(beta1),
         float(beta2), float(eps), float(decay), float(clipvalue), int(step), stream_ptr())


def maxnorm(param, max_norm):
    require_cuda(param)
    rows, cols = param.shape[0], param.numel() // param.shape[0]
    call("gr_maxnorm_f32", ptr(param), rows, cols, float(max_norm), stream_ptr())


def dropout_mask(shape, p, seed, offset, device):
    out = torch.empty(shape, dtype=torch.float32, device=device)
    call("gr_dropout_mask_f32", ptr(out), out.numel(), float(p), int(seed), int(offset), stream_ptr())
    return out


def gaussian_noise(shape, stddev, seed, offset, device):
    out = torch.empty(shape, dtype=torch.float32, device=device)
    call("gr_gaussian_noise_f32", ptr(out), out.numel(), float(stddev), int(seed), int(offset), stream_ptr())
    return out
