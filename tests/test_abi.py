"""The C-ABI library loads on a CPU-only box and exports every symbol include/gr_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "gr_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gr_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_entry_points():
    syms = _header_symbols()
    for must in ("gr_ctc_loss_grad_f32", "gr_ctc_bestpath_ref_f32", "gr_ctc_greedy_f32", "gr_ctc_beam_f32",
                 "gr_lstm_recurrence_fwd_f32", "gr_lstm_recurrence_bwd_f32", "gr_gemm_bf16x3_f32"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    import mgr_b200
    lib = ctypes.CDLL(mgr_b200._lib.LIB_PATH)
    for s in _header_symbols():
        assert hasattr(lib, s), "libgr_b200.so does not export %s" % s
    assert mgr_b200._lib.load().gr_version() >= 100


def test_python_binding_covers_header():
    import mgr_b200
    assert set(_header_symbols()) == set(mgr_b200._lib.exported_symbols())


def test_argument_errors_are_synchronous_without_gpu():
    import mgr_b200
    lib = mgr_b200._lib.load()
    n = ctypes.c_size_t(0)
    assert lib.gr_ctc_workspace_bytes(4, 10, 5, 3, ctypes.byref(n)) == 0 and n.value > 0
    assert lib.gr_ctc_workspace_bytes(0, 10, 5, 3, ctypes.byref(n)) == -1
    assert lib.gr_ctc_loss_grad_f32(None, 0, 1, 1, 2, 0, 1e-8, None, 1, None, None, None, None, None, None, None, 0, None) == -1
    assert b"null" in lib.gr_last_error()


def test_product_refuses_cpu_tensors():
    import pytest
    import torch
    import mgr_b200
    y = torch.full((2, 6, 4), 0.25)
    with pytest.raises(mgr_b200.GrError):
        mgr_b200.ctc_lambda_func([y, torch.zeros(2, 2), torch.full((2, 1), 4), torch.full((2, 1), 1)])
