"""ctypes wrapper of oracle/beam_ref.c (TEST INFRASTRUCTURE): the bit-exact CPU oracle of the
CUDA beam-search kernel.  Build with `make -C oracle` (done by __graft_entry__.build())."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libbeam_ref.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", _HERE])
        _lib = ctypes.CDLL(path)
        _lib.beam_ref.restype = ctypes.c_int
        for n in ("dm_test_expf", "dm_test_logf", "dm_test_log1pf"):
            getattr(_lib, n).restype = ctypes.c_float
            getattr(_lib, n).argtypes = [ctypes.c_float]
        _lib.dm_test_lse.restype = ctypes.c_float
        _lib.dm_test_lse.argtypes = [ctypes.c_float, ctypes.c_float]
    return _lib


def beam_search(probs, seq_len, beam_width=100, top_paths=1, merge_repeated=True, eps=1e-8):
    """probs (T, C) float32.  Returns list of (labels, logp float32) per path."""
    lib = _load()
    p = np.ascontiguousarray(probs, dtype=np.float32)
    T, C = p.shape
    ids = np.empty((top_paths, T), dtype=np.int32)
    lens = np.empty(top_paths, dtype=np.int32)
    logp = np.empty(top_paths, dtype=np.float32)
    rc = lib.beam_ref(p.ctypes.data_as(ctypes.c_void_p), T, C, int(seq_len), ctypes.c_float(eps), int(beam_width),
                      int(top_paths), int(bool(merge_repeated)), ids.ctypes.data_as(ctypes.c_void_p),
                      lens.ctypes.data_as(ctypes.c_void_p), logp.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return [(ids[k, :lens[k]].tolist(), logp[k]) for k in range(top_paths)]


def det(name, *args):
    return getattr(_load(), "dm_test_" + name)(*args)
