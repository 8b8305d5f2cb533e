"""clock64 trace of lstm_tcu.cu (GR_TC_TRACE): per (step, tile) stamps of the TMA / MMA / epilogue roles."""
import os, sys, ctypes
os.environ["GR_TC_TRACE"] = "1"
os.environ["GR_LSTM_IMPL"] = "tcu"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mgr_b200 import ops, _lib
dev = torch.device("cuda:0")
keep = os.environ.get("KEEP", "0") == "1"
shapes = [tuple(int(v) for v in x.split("x")) for x in os.environ.get("SHAPES", "256x500,32x500").split(",")]
T = 120
names = ["tma_poll_done", "tma_issued", "mma_first_full", "mma_commit", "epi_tmem_full", "epi_ld_xpose", "epi_published", "epi_bar2",
         "epi_red", "epi_P_in_regs", "epi_staged", "epi_iter_start", "epi_ld_waited", "epi_ld_arrived", "epi_stage1", "mma_pre_commit"]
for (B, H) in shapes:
    gates = torch.randn(B * T, 8 * H, device=dev) * 0.5
    U = torch.randn(2, H, 4 * H, device=dev) / H ** 0.5
    for it in range(2):
        if it == 1:
            torch.cuda.synchronize(); ops.lstm_workspace(B, H, dev).zero_()
        ops.lstm_recurrence_fwd(gates.clone(), U, B, T, H, keep_cell=keep)
    torch.cuda.synchronize()
    lib = _lib.load(); lib.gr_debug_lstm_tcu_trace_offset.restype = ctypes.c_size_t
    off = lib.gr_debug_lstm_tcu_trace_offset(B, H)
    ws = ops.lstm_workspace(B, H, dev)
    tr = ws[off:off + 160 * 256 * 16 * 8].view(torch.int64).reshape(160, 256, 16).cpu().numpy()
    ncta = int((tr[:, 20, 4] != 0).sum())
    tr = tr[:ncta]
    ntl = 2 if B > 16 else 1
    print("B%d H%d keep=%d: %d CTAs, %d tiles per CTA; cycles relative to epi_tmem_full of the same (step, tile); median over n" % (B, H, keep, ncta, ntl))
    lo, hi = 20, min(200, T * ntl - 4)
    for tau in range(ntl):
        sel = np.arange(lo + tau, hi, ntl)
        st = tr[:, sel, :16].astype(np.int64)
        rel = st - st[:, :, 4:5]
        med = np.median(rel, axis=1)
        print(" tile %d:" % tau)
        for i, nm in enumerate(names):
            print("   %-16s min %7d  med %7d  max %7d" % (nm, med[:, i].min(), np.median(med[:, i]), med[:, i].max()))
        per = np.median(np.diff(tr[:, sel, 4], axis=1), axis=1)
        print("   period (tmem_full -> tmem_full next step): med %d ; epi_red(n) -> tma_poll_done(n+ntl): med %d ; mma phase (first_full -> tmem_full): med %d" % (
            np.median(per), np.median(tr[:, sel[1:], 0] - tr[:, sel[:-1], 8]), np.median(tr[:, sel, 4] - tr[:, sel, 2])))
