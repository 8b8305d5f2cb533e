"""ctypes binding of libgr_b200.so (the C ABI declared in include/gr_b200.h).

The product path has NO fallback: if the CUDA library is missing or a call fails, this module
raises.  Nothing here imports ``oracle``.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgr_b200.so")

GR_OK = 0
CTC_STATUS = {
    1: "Labels length is zero in batch %d",
    2: "Not enough time for target transition sequence in batch %d",
    3: "Saw a non-null label (index >= num_classes - 1) following a null label, batch %d",
    5: "sequence_length(%d) out of range",
}


class GrError(RuntimeError):
    pass


class InvalidArgumentError(ValueError):
    """Raised where the reference would see tf.errors.InvalidArgumentError at session.run."""


_P = c_void_p
_SIGNATURES = {
    "gr_version": ([], c_int),
    "gr_last_error": ([], c_char_p),
    "gr_ctc_workspace_bytes": ([c_int, c_int, c_int, c_int, ctypes.POINTER(c_size_t)], c_int),
    "gr_ctc_loss_grad_f32": ([_P, c_int, c_int, c_int, c_int, c_int, c_float, _P, c_int, _P, _P, _P, _P, _P, _P,
                              _P, c_size_t, _P], c_int),
    "gr_ctc_bestpath_ref_f32": ([_P, c_int, c_int, c_int, c_int, c_double, _P, _P, _P], c_int),
    "gr_ctc_greedy_f32": ([_P, c_int, c_int, c_int, _P, c_float, _P, _P, _P, _P], c_int),
    "gr_ctc_beam_workspace_bytes": ([c_int, c_int, c_int, c_int, ctypes.POINTER(c_size_t)], c_int),
    "gr_ctc_beam_f32": ([_P, c_int, c_int, c_int, _P, c_float, c_int, c_int, c_int, _P, _P, _P, _P, c_size_t, _P],
                        c_int),
    "gr_lstm_workspace_bytes": ([c_int, c_int, ctypes.POINTER(c_size_t)], c_int),
    "gr_lstm_recurrence_fwd_f32": ([_P, _P, c_int, c_int, c_int, _P, _P, _P, c_size_t, _P], c_int),
    "gr_lstm_recurrence_aux_supported": ([c_int, c_int], c_int),
    "gr_lstm_recurrence_grid": ([c_int, c_int], c_int),
    "gr_lstm_recurrence_fwd_aux_f32": ([_P, _P, c_int, c_int, c_int, _P, _P, _P, c_int, c_int, _P, c_size_t, _P], c_int),
    "gr_lstm_recurrence_bwd_f32": ([_P, _P, _P, _P, c_int, c_int, c_int, _P, c_size_t, _P], c_int),
    "gr_gemm_bf16x3_f32": ([_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P],
                           c_int),
    "gr_gemm_a32_f32": ([_P, c_int, c_int, c_int, _P, c_int, c_int, _P, _P, c_int, _P, _P, c_int, c_int, c_int, c_int,
                         c_int, _P], c_int),
    "gr_gemm_a32_dropout_f32": ([_P, c_int, c_int, c_int, _P, c_float, c_int, c_int, _P, _P, c_int, _P, _P, c_int, c_int,
                                 c_int, c_int, c_int, _P], c_int),
    "gr_gemm_simt_f32": ([_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P], c_int),
    "gr_split_bf16_f32": ([_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int, _P], c_int),
    "gr_mask_mul_acc_f32": ([_P, _P, _P, c_int, c_int, c_int, c_int, _P], c_int),
    "gr_dense_softmax_fwd_f32": ([_P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P], c_int),
    "gr_dense_bwd_f32": ([_P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P], c_int),
    "gr_colsum_f32": ([_P, c_int, c_int, c_int, _P, _P], c_int),
    "gr_add_f32": ([_P, _P, _P, c_size_t, _P], c_int),
    "gr_add_into_f32": ([_P, _P, _P, c_size_t, c_int, c_int, _P], c_int),
    "gr_concat2_f32": ([_P, c_int, _P, c_int, _P, c_size_t, _P], c_int),
    "gr_adam_step_f32": ([_P, _P, _P, _P, c_size_t, c_int, c_int, c_float, c_float, c_float, c_float, c_float,
                          c_float, c_float, c_int64, _P], c_int),
    "gr_pack_f32": ([_P, _P, c_int, _P, _P], c_int),
    "gr_adam_flat_f32": ([_P, _P, c_int, _P, _P, _P, c_float, c_float, c_float, c_float, c_float, c_float, c_int64, _P],
                         c_int),
    "gr_maxnorm_f32": ([_P, c_int, c_int, c_float, _P], c_int),
    "gr_dropout_mask_f32": ([_P, c_size_t, c_float, c_uint64, c_uint64, _P], c_int),
    "gr_gaussian_noise_f32": ([_P, c_size_t, c_float, c_uint64, c_uint64, _P], c_int),
}

_lib = None
launch_count = 0  # number of C-ABI compute calls issued (bench.py reports it)


def load():
    """Load the shared library (idempotent).  Raises GrError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GrError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGNATURES)


# ---- optional per-kernel CUDA-event timing (bench.py: roofline of the dominant kernel) --------
_timing = None            # name -> [(start_event, end_event), ...]
_pending_work = None
kernel_timing_shapes = {}  # name -> [algorithmic bytes or flops of each timed launch]


def kernel_timing_begin(names):
    global _timing
    _timing = {n: [] for n in names}
    kernel_timing_shapes.clear()


def kernel_timing_end():
    """Synchronise and return name -> [ms per launch]."""
    global _timing
    import torch
    torch.cuda.synchronize()
    out = {n: [a.elapsed_time(b) for a, b in evs] for n, evs in (_timing or {}).items() if evs}
    _timing = None
    return out


def note_work(amount):
    """Algorithmic bytes/flops of the NEXT call (recorded only while timing is on)."""
    global _pending_work
    _pending_work = amount


# entry points that are timed under another entry point's name (same kernel family)
_TIMING_ALIAS = {"gr_gemm_a32_dropout_f32": "gr_gemm_a32_f32", "gr_lstm_recurrence_fwd_aux_f32": "gr_lstm_recurrence_fwd_f32"}


def call(name, *args):
    global launch_count, _pending_work
    lib = load()
    tname = _TIMING_ALIAS.get(name, name)
    timed = _timing is not None and tname in _timing
    if timed:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(lib, name)(*args)
    if rc != GR_OK:
        raise GrError("%s failed with code %d: %s" % (name, rc, lib.gr_last_error().decode()))
    if timed:
        e1.record()
        _timing[tname].append((e0, e1))
        kernel_timing_shapes.setdefault(tname, []).append(_pending_work or 0)
    _pending_work = None
    launch_count += 1
    return rc


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise GrError("non-contiguous tensor passed to the C ABI (shape %s, strides %s)" % (tuple(t.shape), t.stride()))
    return c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise GrError("mgr_b200 ops need CUDA tensors (no CPU fallback); got a %s tensor" % t.device)
