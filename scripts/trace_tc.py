import os, sys, ctypes
os.environ["GR_TC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mgr_b200 import ops, _lib
dev = torch.device("cuda:0")
keep = os.environ.get("KEEP", "0") == "1"
shapes = [(256, 100, 500), (256, 100, 300)]
for (B, T, H) in shapes:
    gates = torch.randn(B * T, 8 * H, device=dev) * 0.5
    U = torch.randn(2, H, 4 * H, device=dev) / H ** 0.5
    for it in range(2):
        if it == 1:
            torch.cuda.synchronize(); ops.lstm_workspace(B, H, dev).zero_()
        ops.lstm_recurrence_fwd(gates.clone(), U, B, T, H, keep_cell=keep)
    torch.cuda.synchronize()
    lib = _lib.load(); lib.gr_debug_lstm_tc_trace_offset.restype = ctypes.c_size_t
    off = lib.gr_debug_lstm_tc_trace_offset(B, H)
    ws = ops.lstm_workspace(B, H, dev)
    tr = ws[off:off + 160 * 128 * 16 * 8].view(torch.int64).reshape(160, 128, 16).cpu().numpy()
    ncta = int((tr[:, 20, 0] != 0).sum())
    tr = tr[:ncta]
    print("B%d H%d keep=%d: %d CTAs; cycles, median over steps 10..90" % (B, H, keep, ncta))
    names = ["poll_done", "tma_issued", "mma_first_full", "mma_committed", "epi_tmem_full", "epi_ld_done", "epi_publish_done", "bar_sync_done", "red_done", "full c0/1", "full c2/3", "full c4/5", "full c6/7"]
    st = tr[:, 10:90, :13]
    rel = st - st[:, :, 0:1]
    med = np.median(rel, axis=1)           # (cta, slot)
    for i, n in enumerate(names):
        print("  %-18s min %6d  med %6d  max %6d" % (n, med[:, i].min(), np.median(med[:, i]), med[:, i].max()))
    step_len = np.median(np.diff(tr[:, 10:90, 0], axis=1), axis=1)
    wait = np.median(tr[:, 11:90, 0] - tr[:, 10:89, 8], axis=1)
    print("  step length: min %d med %d max %d ; red_done(s-1)->poll_done(s): min %d med %d max %d" % (
        step_len.min(), np.median(step_len), step_len.max(), wait.min(), np.median(wait), wait.max()))
    # cross-CTA skew from globaltimer (ns): per step, spread of arrive times and of poll_done times
    ga = tr[:, 10:90, 10].astype(np.float64); gp = tr[:, 10:90, 9].astype(np.float64)
    print("  globaltimer: arrive spread (max-min over CTAs) med %.0f ns; last arrive -> first poll_done(s+1) med %.0f ns; -> last poll_done med %.0f ns" % (
        np.median(ga.max(0) - ga.min(0)), np.median(gp[:, 1:].min(0) - ga[:, :-1].max(0)), np.median(gp[:, 1:].max(0) - ga[:, :-1].max(0))))
    slow = np.argsort(-med[:, 7])[:6]
    print("  slowest CTAs by bar_sync_done:", [(int(c), int(med[c, 3]), int(med[c, 6]), int(med[c, 7])) for c in slow])
