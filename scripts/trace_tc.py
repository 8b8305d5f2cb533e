import os, sys, ctypes
os.environ["GR_TC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mgr_b200 import ops, _lib
dev = torch.device("cuda:0")
for (B, T, H) in [(256, 100, 500), (256, 100, 300), (32, 100, 500)]:
    gates = torch.randn(B * T, 8 * H, device=dev) * 0.5
    U = torch.randn(2, H, 4 * H, device=dev) / H ** 0.5
    for _ in range(2):
        ops.lstm_recurrence_fwd(gates.clone(), U, B, T, H, keep_cell=False)
    torch.cuda.synchronize()
    lib = _lib.load(); lib.gr_debug_lstm_tc_trace_offset.restype = ctypes.c_size_t
    off = lib.gr_debug_lstm_tc_trace_offset(B, H)
    ws = ops.lstm_workspace(B, H, dev)
    tr = ws[off:off + 128 * 8 * 8].view(torch.int64).reshape(128, 8).cpu().numpy()
    print("B%d H%d: per-step deltas (cycles), median over steps 10..90" % (B, H))
    names = ["poll_done", "tma_issued", "mma_first_full", "mma_committed", "epi_tmem_full", "epi_ld_done", "epi_math_done", "epi_barrier_done"]
    st = tr[10:90]
    base = st[:, 0:1]
    rel = st - base
    print("  rel to poll_done:", {n: int(np.median(rel[:, i])) for i, n in enumerate(names)})
    step_len = np.diff(tr[10:90, 0])
    print("  step length median %d cycles; barrier_done(s-1)->poll_done(s): %d" % (np.median(step_len), np.median(tr[11:90, 0] - tr[10:89, 7])))
