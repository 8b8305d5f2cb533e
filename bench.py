#!/usr/bin/env python
"""bench.py -- fusion BLSTM-CTC training throughput (BASELINE.json config 3) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W     # CPU restatement of the reference path

A "step" is one full training step of multimodal_fusion/multimodal.py on synthetic data:
frozen speech+skeletal towers (dropout/noise active, learning phase 1) -> concat -> BLSTM(100) ->
Dropout -> Dense(22) -> fused softmax+CTC loss/grad -> BPTT through the fusion BLSTM -> one flat
NCCL all-reduce of the 1 365 222 trainable gradients -> Adam(clipvalue)+maxnorm.  Global batch 256,
T=1000, sharded over the ranks (strong scaling, as BASELINE.json states the config).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GLOBAL_BATCH = 256
T_FRAMES = 1000
NB_CLASSES = 22
LMAX = 35
METRIC = "fusion BLSTM-CTC train seq/s"


def _traffic_entry(kernel):
    """Entry of `kernel` in the newest committed `ncu --set full` summary (profiles/rNN_traffic.json)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        try:
            d = json.load(open(path)).get(kernel)
        except Exception:
            d = None
        if d is not None:
            return d, name
    return None, None


def read_traffic_shape(kernel):
    d, name = _traffic_entry(kernel)
    return None if d is None else "%s (profiles/%s)" % (d["shape"], name)


def read_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full`
    capture, or None."""
    d, _ = _traffic_entry(kernel)
    return None if d is None else d["dram_bytes"]


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self, name):
        """Wall-clock marks: samples between 'begin' and 'end' belong to the timed region."""
        self.marks = getattr(self, "marks", {})
        self.marks[name] = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        import datetime
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        marks = getattr(self, "marks", {})
        t0, t1 = marks.get("begin"), marks.get("end")
        rows = []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), float(f[3]), f[4:8]))
            except ValueError:
                continue
        inside = [r for r in rows if t0 is not None and t1 is not None and t0 - 0.05 <= r[0] <= t1 + 0.05]
        window = "timed region"
        if len(inside) < 2:   # region shorter than the sampling period: fall back to every sample under load
            inside = [r for r in rows if r[3] > 250.0] or rows
            window = "whole GPU arm (timed region shorter than 2 samples)"
        reasons = set()
        for r in inside:
            for n, v in zip(names, r[4]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm = [r[1] for r in inside]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max([r[2] for r in inside]) if inside else None,
                "power_w_max": max([r[3] for r in inside]) if inside else None, "samples": len(inside), "window": window,
                "reasons": sorted(reasons)}


def synth_batch(rows_lo, rows_hi, T):
    """Rows [lo,hi) of the GLOBAL synthetic batch (seeds 2001/2002/2003: SURVEY.md 8d config 3)."""
    import torch
    ga = torch.Generator().manual_seed(2001)
    gs = torch.Generator().manual_seed(2002)
    xa = torch.randn(GLOBAL_BATCH, T, 39, generator=ga)[rows_lo:rows_hi].contiguous()
    xs = torch.randn(GLOBAL_BATCH, T, 20, generator=gs)[rows_lo:rows_hi].contiguous()
    rng = np.random.default_rng(2003)
    labels = -np.ones((GLOBAL_BATCH, LMAX), dtype=np.float32)
    ll = np.zeros((GLOBAL_BATCH, 1), dtype=np.int64)
    for b in range(GLOBAL_BATCH):
        L = int(rng.integers(1, LMAX + 1))
        labels[b, :L] = rng.integers(0, NB_CLASSES - 1, size=L)
        ll[b, 0] = L
    il = np.full((GLOBAL_BATCH, 1), T - 2, dtype=np.int64)
    return xa, xs, torch.from_numpy(labels[rows_lo:rows_hi]), torch.from_numpy(il[rows_lo:rows_hi]), \
        torch.from_numpy(ll[rows_lo:rows_hi])


# --------------------------------------------------------------------------------------- CPU arm
def cpu_reference_step_time(n_seq, T, steps=1, warmup=0):
    """The restated reference path (oracle/lstm_ref.py, torch CPU fp32, all host threads):
    fusion forward (towers + fusion BLSTM + Dense + softmax), ctc loss, backward through the
    fusion BLSTM + Dense.  Returns (seconds per step, threads)."""
    import torch
    from oracle import lstm_ref
    torch.set_num_threads(os.cpu_count() or 1)
    rng = np.random.default_rng(47)
    t32 = lambda ws: [torch.tensor(w) for w in ws]
    sp1, sp2 = t32(lstm_ref.init_blstm_weights(rng, 39, 500)), t32(lstm_ref.init_blstm_weights(rng, 1000, 500))
    sk1, sk2 = t32(lstm_ref.init_blstm_weights(rng, 20, 300)), t32(lstm_ref.init_blstm_weights(rng, 600, 300))
    fu = [w.requires_grad_(True) for w in t32(lstm_ref.init_blstm_weights(rng, 1600, 100))]
    fd = [w.requires_grad_(True) for w in t32(lstm_ref.init_dense_weights(rng, 200, NB_CLASSES))]
    xa, xs, labels, il, ll = synth_batch(0, n_seq, T)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        noise = torch.randn_like(xa) * 0.5
        masks = {"fu_f": (torch.rand(4, n_seq, 1600) > 0.5).float() * 2, "fu_b": (torch.rand(4, n_seq, 1600) > 0.5).float() * 2}
        p, a, _ = lstm_ref.fusion_forward(xa, xs, sp1, sp2, sk1, sk2, fu, fd, noise_a=noise, masks=masks,
                                          drop_mask=(torch.rand(n_seq, T, 200) > 0.5).float() * 2)
        loss = lstm_ref.torch_ctc_lambda(p, labels, il, ll).mean()
        loss.backward()
        for w in fu + fd:
            w.grad = None
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return float(np.mean(times)), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_seq = args.ref_batch            # bounded sample (32 sequences, ~4 s of host work per step): the SAME sample the GPU arm's
    #                                   `cpu_baseline` object times
    sec, threads = cpu_reference_step_time(n_seq, T_FRAMES, steps=args.steps, warmup=min(args.warmup, 1))
    v = n_seq / sec
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "seq/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "fusion BLSTM-CTC training step (multimodal.py), T=1000, C=22; CPU sample of "
                                   "%d sequences per step (same sample as the GPU arm's cpu_baseline; 1 warm-up step)" % n_seq,
                       "global_batch": GLOBAL_BATCH, "seq_len": T_FRAMES},
            "cpu_baseline": {"value": v, "unit": "seq/s", "cores": threads, "kind": "port",
                             "sample": "%d sequences x T=%d per step, %d steps, torch-CPU fp32 restatement "
                                       "(oracle/lstm_ref.py)" % (n_seq, T_FRAMES, args.steps)},
            "e2e": {"value": v, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- GPU arm
def ctc_microbench(dev, peak_gbs):
    """BASELINE config 4 point: B=1024, T=1000 (+2 dropped), L=40, C=22; CUDA events, L2-exceeding
    working set (lattice 360 MB + probs/grad 180 MB)."""
    import torch
    from mgr_b200 import ops
    B, T, C, L = 1024, 1002, 22, 40
    g = torch.Generator().manual_seed(3001)
    probs = torch.softmax(torch.randn(B, T, C, generator=g) * 2, -1).to(dev)
    rng = np.random.default_rng(3002)
    labels = torch.tensor(rng.integers(0, C - 1, size=(B, L)), dtype=torch.int32, device=dev)
    ll = torch.full((B,), L, dtype=torch.int32, device=dev)
    il = torch.full((B,), T - 2, dtype=torch.int32, device=dev)
    for _ in range(3):
        ops.ctc_loss_grad(probs, labels, ll, il, False)
    torch.cuda.synchronize()
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        ops.ctc_loss_grad(probs, labels, ll, il, False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    frames = B * (T - 2)
    bytes_per_frame = 8 * C + 8 * (2 * L + 1) + (4 * L + 16) / (T - 2)
    achieved = frames * bytes_per_frame / (ms * 1e-3) / 1e9
    return {"workload": "ctc loss+grad B=1024 T=1000 L=40 C=22 (config 4)", "ms": ms,
            "frames_per_s": frames / (ms * 1e-3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                         "frac": achieved / peak_gbs, "traffic": read_traffic("gr_ctc_loss_grad_f32"),
                         "traffic_capture": read_traffic_shape("gr_ctc_loss_grad_f32"),
                         "algorithmic_bytes_per_frame": bytes_per_frame}}


def ctc_sweep(dev, peak_gbs, full=False):
    """BASELINE config 4 as written: B=1024, T in {100..2000}, label length up to 40 (capped at T/2), C=22 and 44,
    full and ragged input lengths.  Default: L in {10, 40} (56 shapes); `--ctc-sweep`: L in {1,10,20,40}."""
    import torch
    from mgr_b200 import ops
    B = 1024
    out = []
    rng = np.random.default_rng(3002)
    for C, T in [(c, t) for c in (22, 44) for t in (100, 200, 400, 800, 1000, 1600, 2000)]:
        g = torch.Generator(device=dev).manual_seed(3001 + T + C)
        probs = torch.softmax(torch.randn(B, T + 2, C, generator=g, device=dev) * 2, -1)
        for L in ((1, 10, 20, 40) if full else (10, 40)):
            L = min(L, T // 2)
            labels = torch.tensor(rng.integers(0, C - 1, size=(B, L)), dtype=torch.int32, device=dev)
            ll = torch.full((B,), L, dtype=torch.int32, device=dev)
            for ragged in (False, True):
                il_np = rng.integers(max(T // 2, L), T + 1, size=B) if ragged else np.full(B, T)
                il = torch.tensor(il_np, dtype=torch.int32, device=dev)
                for _ in range(2):
                    ops.ctc_loss_grad(probs, labels, ll, il, False)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 5
                e0.record()
                for _ in range(n):
                    ops.ctc_loss_grad(probs, labels, ll, il, False)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                frames = int(il_np.sum())
                bpf = 8 * C + 8 * (2 * L + 1)
                ach = frames * bpf / (ms * 1e-3) / 1e9
                out.append({"C": C, "T": T, "L": L, "ragged": ragged, "ms": round(ms, 4),
                            "Gframes_per_s": round(frames / (ms * 1e-3) / 1e9, 3),
                            "achieved_GBps": round(ach, 1), "frac": round(ach / peak_gbs, 3)})
        del probs
    return out


def decode_microbench(dev, peak_gbs):
    """BASELINE config 5: N=512, T=1000, C=22 'peaky' probabilities; thresholded best path
    (sequence_decoding.py semantics), TF greedy and beam-width-100 decode.  CUDA events."""
    import torch
    from mgr_b200 import ops
    N, T, C = 512, 1002, 22
    g = torch.Generator().manual_seed(4001)
    seg = torch.randint(0, C, (N, T // 20 + 2), generator=g)
    track = seg.repeat_interleave(20, dim=1)[:, :T]
    logits = torch.randn(N, T, C, generator=g)
    logits.scatter_add_(2, track.unsqueeze(-1), torch.full((N, T, 1), 4.0))
    probs = torch.softmax(logits, -1).to(dev)
    # four copies (180 MB > 126 MB L2), used round-robin, so that no iteration reads its input from L2
    copies = [probs] + [probs.clone() for _ in range(3)]
    it = [0]

    def nxt():
        it[0] += 1
        return copies[it[0] % len(copies)]

    def timeit(fn, n):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    frames = N * (T - 2)
    bytes_bp = 4.0 * N * T * C + 4.0 * N * T
    ms_bp = timeit(lambda: ops.bestpath_ref(nxt(), 0.5), 20)
    ms_gr = timeit(lambda: ops.greedy(nxt()), 20)
    ms_bm = timeit(lambda: ops.beam(nxt(), beam_width=100), 3)
    return {"workload": "decode N=512 T=1000 C=22 (config 5)",
            "bestpath_ref": {"ms": ms_bp, "frames_per_s": frames / (ms_bp * 1e-3),
                             "roofline": {"bound": "hbm", "achieved": bytes_bp / (ms_bp * 1e-3) / 1e9, "peak": peak_gbs,
                                          "unit": "GB/s", "frac": bytes_bp / (ms_bp * 1e-3) / 1e9 / peak_gbs}},
            "greedy": {"ms": ms_gr, "frames_per_s": frames / (ms_gr * 1e-3)},
            "beam100": {"ms": ms_bm, "frames_per_s": N * T / (ms_bm * 1e-3)}}


def unimodal_leg(dev, hbm_peak, tc_peak, kind, steps=4, cpu=True):
    """BASELINE config 2 (skeletal BLSTM-CTC TRAINING step, B=64, T=800: skeletal_lstm_ctc.py:296-394) and config 1
    (speech BLSTM-CTC forward + ctc_batch_cost, B=16, T=400: speech_lstm_ctc_words.py) on one GPU, CUDA events; config 1
    is the reference's CPU-runnable case, so the restated CPU path (oracle/lstm_ref.py) is timed beside it."""
    import torch
    import mgr_b200 as mgr
    from mgr_b200 import _lib
    names = ["gr_lstm_recurrence_fwd_f32", "gr_lstm_recurrence_bwd_f32", "gr_gemm_a32_f32", "gr_ctc_loss_grad_f32",
             "gr_split_bf16_f32"]
    if kind == "skeletal_train":
        B, T, F, C, Lmax = 64, 800, 20, 22, 28
        net = mgr.SkeletalNet().to(dev)
        train = True
    else:
        B, T, F, C, Lmax = 16, 400, 39, 44, 20
        net = mgr.SpeechNet().to(dev)
        train = False
    g = torch.Generator().manual_seed(1002 if train else 1001)
    x = torch.randn(B, T, F, generator=g).to(dev)
    rng = np.random.default_rng(1003)
    labels = -np.ones((B, Lmax), dtype=np.float32)
    ll = np.zeros((B, 1), dtype=np.int64)
    for b in range(B):
        L = int(rng.integers(1, Lmax + 1))
        labels[b, :L] = rng.integers(0, C - 1, size=L)
        ll[b, 0] = L
    il = np.full((B, 1), T - 2, dtype=np.int64)
    lab_t, il_t, ll_t = torch.tensor(labels), torch.tensor(il), torch.tensor(ll)
    params = [p for p in net.parameters()]
    opt = mgr.KerasAdam(params, lr=1e-4, clipvalue=0.5, decay=1e-5,
                        maxnorm_params=[net.blstm_1.kernel, net.blstm_2.kernel], max_norm=3.0) if train else None
    step_no = [0]

    def step():
        if train:
            reg = net.sample_regularisers(B, T, seed=77, step=step_no[0], device=dev)
            step_no[0] += 1
            _, logits = net(x, reg)
            loss = mgr.softmax_ctc(logits, lab_t, il_t, ll_t, check=False)
            grads = torch.autograd.grad(loss.mean(), params)
            opt.step(list(grads))
        else:
            with torch.no_grad():
                y_pred, _ = net(x, None)
                loss = mgr.ctc_lambda_func([y_pred, lab_t, il_t, ll_t])
        return loss

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # per-kernel durations / roofline from a pass with ONE launch per recurrence (the timed steps above run the training
    # recurrences as two concurrent half-batch launches when they fit, layers._halves: event times would overlap)
    split_env = os.environ.get("GR_TRAIN_SPLIT")
    os.environ["GR_TRAIN_SPLIT"] = "0"
    step()
    torch.cuda.synchronize()
    _lib.kernel_timing_begin(names)
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    kt = _lib.kernel_timing_end()
    if split_env is None:
        os.environ.pop("GR_TRAIN_SPLIT")
    else:
        os.environ["GR_TRAIN_SPLIT"] = split_env
    per_kernel = {k: {"ms_per_step": round(sum(v) / steps, 3), "launches_per_step": len(v) / steps} for k, v in kt.items()}
    dom = max(per_kernel.items(), key=lambda kv: kv[1]["ms_per_step"])[0]
    calls = _lib.kernel_timing_shapes.get(dom, [])
    work = sum(calls) / max(1, len(calls))
    avg_ms = sum(kt[dom]) / len(kt[dom])
    if dom == "gr_gemm_a32_f32":
        roof = {"kernel": dom, "bound": "tensor", "achieved": work / (avg_ms * 1e-3) / 1e12, "peak": tc_peak, "unit": "TFLOP/s"}
    else:
        roof = {"kernel": dom, "bound": "hbm", "achieved": work / (avg_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["avg_launch_ms"] = avg_ms
    roof["share_of_step"] = per_kernel[dom]["ms_per_step"] / max(ms, sum(v["ms_per_step"] for v in per_kernel.values()))
    roof["note"] = "kernel durations from steps with one launch per recurrence (GR_TRAIN_SPLIT=0); ms_per_step is the default schedule" 
    roof["traffic"] = None
    out = {"workload": ("skeletal BLSTM-CTC training step (skeletal_lstm_ctc.py), B=64 T=800 F=20 H=300 C=22, regularisers on"
                        if train else "speech BLSTM-CTC forward + ctc_batch_cost (speech_lstm_ctc_words.py), B=16 T=400 F=39 "
                                      "H=500 C=44, learning phase 0"),
           "ms_per_step": ms, "seq_per_s": B / (ms * 1e-3), "loss_mean": float(loss.detach().mean()), "kernels": per_kernel,
           "roofline": roof}
    if cpu and not train:
        from oracle import lstm_ref
        torch.set_num_threads(os.cpu_count() or 1)
        rngw = np.random.default_rng(47)
        t32 = lambda ws: [torch.tensor(w) for w in ws]
        w1, w2 = t32(lstm_ref.init_blstm_weights(rngw, F, 500)), t32(lstm_ref.init_blstm_weights(rngw, 1000, 500))
        wd = t32(lstm_ref.init_dense_weights(rngw, 1000, C))
        xc = x.cpu()
        t0 = time.perf_counter()
        with torch.no_grad():
            pc, _ = lstm_ref.unimodal_forward(xc, w1, w2, wd)[:2]
            lstm_ref.torch_ctc_lambda(pc, lab_t, il_t, ll_t)
        sec = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": B / sec, "unit": "seq/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": "the whole config (16 sequences x T=400), 1 pass, torch-CPU fp32 restatement, %.1f s" % sec}
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import mgr_b200 as mgr
    from mgr_b200 import parallel, _lib

    rank, world, local = parallel.init_from_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    hbm_peak, tc_peak, peak_kind = read_peaks()
    lo, hi = parallel.shard_rows(GLOBAL_BATCH, rank, world)
    B = hi - lo
    T = args.seq_len

    if os.environ.get("GR_MAIN_PRIO"):   # experiment: the fusion layer's stream above the towers' streams
        torch.cuda.set_stream(torch.cuda.Stream(priority=int(os.environ["GR_MAIN_PRIO"])))
    model = mgr.FusionNet().to(dev)
    opt = mgr.fusion_optimizer(model)
    bucket = parallel.FlatGradBucket(model.trainable_parameters())
    xa_h, xs_h, lab_h, il_h, ll_h = [t.pin_memory() for t in synth_batch(lo, hi, T)]
    xa_d, xs_d = xa_h.to(dev), xs_h.to(dev)
    h2d_bytes = sum(t.numel() * t.element_size() for t in (xa_h, xs_h, lab_h, il_h, ll_h))

    def reduce_grads(grads):
        bucket.pack(grads)
        return bucket.all_reduce()

    # FusionTrainer = the reference's training step (forward, CTC objective, all-reduce, Adam/clip/maxnorm) with the
    # frozen towers of the NEXT batch enqueued beside the fusion layer of this one (GR_PIPELINE=0: strictly serial)
    trainer = mgr.FusionTrainer(model, opt, seed=1234 + rank, global_batch=GLOBAL_BATCH, grad_hook=reduce_grads)
    # GR_PIPELINE: 1 (default at every N) = towers one batch ahead, the all-reduce of step n ordered BEHIND the
    # prefetched towers of batch n+1 (FusionTrainer.hook_after_towers: a collective never shares the GPU with the
    # spinning cooperative recurrence kernels -- the combination that did not finish in two round-1 runs; the next
    # fusion step needs those towers anyway, so nothing on the critical path waits longer);  0 = strictly serial;
    # 3 = pipeline with the all-reduce free to overlap the towers (round-1 behaviour, for investigation only).
    pmode = os.environ.get("GR_PIPELINE", "1")
    pipeline = pmode in ("1", "2", "3")
    trainer.hook_after_towers = pmode != "3"
    # towers `depth` batches ahead (FusionTrainer's prefetch queue).  Measured at N = 1: depth 2 / 3 = 39.95-40.23 ms per
    # step against 39.45-39.9 for depth 1 -- a second set of independent tower kernels does not pack any better, so the
    # default stays 1 (GR_PREFETCH_DEPTH for experiments)
    depth = int(os.environ.get("GR_PREFETCH_DEPTH", "1")) if pipeline else 0

    def train_step(xa, xs, lab, il, ll, nxt=None, nxt_ready=None):
        return trainer.step((xa, xs, lab, il, ll), next_inputs=nxt if pipeline else None, next_ready=nxt_ready)

    lab_d, il_d, ll_d = lab_h.to(dev), il_h.to(dev), ll_h.to(dev)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident arm
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ahead = [(xa_d, xs_d)] * max(depth, 1)
    for _ in range(max(args.warmup, args.min_warmup)):
        train_step(xa_d, xs_d, lab_d, il_d, ll_d, ahead)
    launches0 = _lib.launch_count
    _lib.kernel_timing_begin(["gr_lstm_recurrence_fwd_f32", "gr_lstm_recurrence_bwd_f32", "gr_gemm_bf16x3_f32",
                              "gr_ctc_loss_grad_f32", "gr_split_bf16_f32", "gr_gemm_a32_f32"])
    sampler.mark("begin")
    # every timed step trains one batch and enqueues the towers of the next one (K tower passes + K fusion passes)
    total_ms = timed(lambda: train_step(xa_d, xs_d, lab_d, il_d, ll_d, ahead), args.steps)
    sampler.mark("end")
    ktimes = _lib.kernel_timing_end()
    launches = _lib.launch_count - launches0
    ms_per_step = total_ms / args.steps
    value = GLOBAL_BATCH / (ms_per_step * 1e-3)

    # ---- end-to-end arm: pinned host inputs -> H2D every step, loss read back every step
    losses = []

    copy_stream = torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(copy_stream):
            ts = tuple(t.to(dev, non_blocking=True) for t in (xa_h, xs_h, lab_h, il_h, ll_h))
            ev = copy_stream.record_event()
        return ts, ev

    trainer.close()                      # the device-resident arm's prefetches are for other tensors
    staged = [h2d() for _ in range(max(depth, 1))]

    def e2e_step():
        # one H2D copy of a full batch from pinned memory (on a copy stream) and one D2H read of the loss per step;
        # with the pipeline the copy made in step n is the batch of step n+depth, whose towers start as soon as it lands
        (cur, ev) = staged.pop(0)
        torch.cuda.current_stream().wait_event(ev)
        for t in cur:
            t.record_stream(torch.cuda.current_stream())
        if pipeline:
            staged.append(h2d())
            nxt = [(b_[0], b_[1]) for b_, _ in staged]
            loss = train_step(*cur, nxt=nxt, nxt_ready=[e_ for _, e_ in staged])
        else:
            loss = train_step(*cur)
            staged.append(h2d())
        losses.append(loss.cpu())

    if args.skip_e2e:
        if rank == 0:
            sampler.stop()
        return
    e2e_step()
    e2e_ms = timed(e2e_step, args.steps) / args.steps
    e2e_value = GLOBAL_BATCH / (e2e_ms * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    # ---- kernel pass: in the timed region up to four tower streams and the fusion layer overlap, so a launch's event
    # time includes time-sharing with other kernels (per-kernel sums exceed the step).  The per-kernel durations the
    # roofline is computed from are therefore taken live, right here, from SERIAL steps: one stream, whole-batch
    # towers, no pipeline -- each kernel alone on the GPU, as in the committed ncu launch list.
    overlapped = {k: {"ms_per_step": sum(v) / args.steps, "launches_per_step": len(v) / args.steps,
                      "avg_ms": sum(v) / max(1, len(v))} for k, v in ktimes.items()}
    saved_env = {k: os.environ.get(k) for k in ("GR_TOWER_STREAMS", "GR_TOWER_SPLIT")}
    os.environ["GR_TOWER_STREAMS"], os.environ["GR_TOWER_SPLIT"] = "0", "0"
    serial_steps = 2

    def serial_step():
        trainer.step((xa_d, xs_d, lab_d, il_d, ll_d))

    serial_step()
    _lib.kernel_timing_begin(["gr_lstm_recurrence_fwd_f32", "gr_lstm_recurrence_bwd_f32", "gr_gemm_bf16x3_f32",
                              "gr_ctc_loss_grad_f32", "gr_split_bf16_f32", "gr_gemm_a32_f32"])
    serial_ms = timed(serial_step, serial_steps) / serial_steps
    ktimes = _lib.kernel_timing_end()
    for k, v in saved_env.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (share of the serial step, CUDA events)
    per_kernel = {k: {"ms_per_step": sum(v) / serial_steps, "launches_per_step": len(v) / serial_steps,
                      "avg_ms": sum(v) / max(1, len(v))} for k, v in ktimes.items()}
    dom = max(per_kernel.items(), key=lambda kv: kv[1]["ms_per_step"])[0] if per_kernel else None
    roofline = None
    if dom == "gr_lstm_recurrence_fwd_f32" or dom == "gr_lstm_recurrence_bwd_f32":
        # algorithmic bytes of the recurrence per launch, both directions:
        # read 4H pre-activations + write h (+ 4H gates + c when kept for BPTT) per (b, t, unit)
        calls = _lib.kernel_timing_shapes.get(dom, [])
        # The entry point is launched on several layer shapes per step (speech / skeletal towers, fusion layer).  The
        # roofline is stated for ONE launch shape -- the group of launches with equal algorithmic bytes that takes the most
        # time (the speech-tower layers; the shape of the committed ncu capture behind `traffic`) -- and the average over
        # all launches of the entry point is kept beside it.
        groups = {}
        for b_, t_ in zip(calls, ktimes[dom]):
            groups.setdefault(b_, []).append(t_)
        gb, gt = max(groups.items(), key=lambda kv: sum(kv[1]))
        avg_ms = sum(gt) / len(gt)
        ach = gb / (avg_ms * 1e-3) / 1e9
        all_bytes = sum(c for c in calls) / max(1, len(calls))
        all_ms = per_kernel[dom]["avg_ms"]
        roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": read_traffic(dom), "traffic_capture": read_traffic_shape(dom),
                    "peak_kind": peak_kind,
                    "avg_launch_ms": avg_ms, "launches_of_this_shape_per_step": len(gt) / serial_steps,
                    "algorithmic_bytes_per_launch": gb,
                    "all_launches": {"avg_launch_ms": all_ms, "achieved": all_bytes / (all_ms * 1e-3) / 1e9,
                                     "frac": all_bytes / (all_ms * 1e-3) / 1e9 / hbm_peak,
                                     "launches_per_step": per_kernel[dom]["launches_per_step"]},
                    "share_of_step": per_kernel[dom]["ms_per_step"] / serial_ms,
                    "serial_step_ms": serial_ms,
                    "note": "durations from serial steps run right after the timed region (kernel alone on the GPU); "
                            "algorithmic bytes = 20 B (inference) / 40 B (training) per (b,t,unit,dir); the kernel is "
                            "bound by T serial steps (latency), not by HBM: see DESIGN.md 4.3"}
    elif dom in ("gr_gemm_bf16x3_f32", "gr_gemm_a32_f32"):
        calls = _lib.kernel_timing_shapes.get(dom, [])
        flops = sum(calls) / max(1, len(calls))
        avg_ms = per_kernel[dom]["avg_ms"]
        ach = flops / (avg_ms * 1e-3) / 1e12
        roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": tc_peak, "unit": "TFLOP/s",
                    "frac": ach / tc_peak, "traffic": read_traffic(dom), "traffic_capture": read_traffic_shape(dom),
                    "peak_kind": peak_kind, "avg_launch_ms": avg_ms,
                    "share_of_step": per_kernel[dom]["ms_per_step"] / serial_ms, "serial_step_ms": serial_ms,
                    "note": "durations from serial steps run right after the timed region; algorithmic flops 2MNK "
                            "(bf16x3 executes 3x that on the tensor pipe)"}
    # ---- secondary fusion line at the reference's own maxlen (multimodal.py:222): T=1900, same schedule, 3 steps
    t1900 = None
    if not args.skip_ctc and T != 1900 and world == 1:
        xa9, xs9, lab9, il9, ll9 = [t.to(dev) for t in synth_batch(lo, hi, 1900)]
        tr9 = mgr.FusionTrainer(model, opt, seed=4321 + rank, global_batch=GLOBAL_BATCH, grad_hook=reduce_grads)
        for _ in range(2):
            tr9.step((xa9, xs9, lab9, il9, ll9), next_inputs=(xa9, xs9) if pipeline else None)
        ms9 = timed(lambda: tr9.step((xa9, xs9, lab9, il9, ll9), next_inputs=(xa9, xs9) if pipeline else None), 3) / 3
        tr9.close()
        t1900 = {"seq_len": 1900, "ms_per_step": ms9, "value": GLOBAL_BATCH / (ms9 * 1e-3), "unit": "seq/s",
                 "frames_per_s": GLOBAL_BATCH * 1900 / (ms9 * 1e-3)}
        del xa9, xs9, tr9
        torch.cuda.empty_cache()
    ctc = ctc_microbench(dev, hbm_peak) if not args.skip_ctc else None
    if ctc is not None:
        ctc["sweep"] = ctc_sweep(dev, hbm_peak, full=args.ctc_sweep)
    decode = decode_microbench(dev, hbm_peak) if not args.skip_ctc else None
    config2 = unimodal_leg(dev, hbm_peak, tc_peak, "skeletal_train") if not args.skip_ctc else None
    config1 = unimodal_leg(dev, hbm_peak, tc_peak, "speech_fwd", cpu=not args.skip_cpu) if not args.skip_ctc else None
    cpu = None
    if not args.skip_cpu:
        sec, threads = cpu_reference_step_time(args.ref_batch, T, steps=1, warmup=1)
        cpu = {"value": args.ref_batch / sec, "unit": "seq/s", "cores": threads, "kind": "port",
               "sample": "%d sequences x T=%d, 1 warm-up + 1 timed step, torch-CPU fp32 restatement (oracle/lstm_ref.py), %.1f s"
                         % (args.ref_batch, T, sec)}
    line = {"metric": METRIC, "value": value, "unit": "seq/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, args.min_warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "fusion BLSTM-CTC training step (multimodal.py topology: speech 39->2xBLSTM500, "
                                   "skeletal 20->2xBLSTM300 frozen, fusion BLSTM100 + Dense22 + CTC trained), "
                                   "regularisers on (Philox)", "global_batch": GLOBAL_BATCH, "per_gpu_batch": B,
                       "seq_len": T, "classes": NB_CLASSES, "parallelism": "dp%d" % world,
                       "projection_arithmetic": "bf16x3 split on tcgen05 (fp32-faithful)",
                       "schedule": ("frozen towers of batch n+1 run beside the fusion layer of batch n "
                                    "(FusionTrainer; every timed step = one tower pass + one fusion pass + optimiser)"
                                    if pipeline else "serial (GR_PIPELINE=0)"),
                       "l2": "per-step activation working set (>10 GB) exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "seq/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": B * 4,
                    "note": "inputs: pinned host -> device on a copy stream, one batch ahead; loss: blocking "
                            "device -> host read every step"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": per_kernel,
            "kernels_in_timed_region": overlapped,
            "fusion_T1900": t1900, "config1_speech_fwd_loss": config1, "config2_skeletal_train": config2,
            "ctc": ctc, "decode": decode, "cpu_baseline": cpu, "loss_mean": float(torch.cat(losses).mean())}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def start_watchdog(seconds):
    """A hung collective or kernel must not hold the box until the caller's limit: after `seconds` the process
    prints one line to stderr and exits (which tears the CUDA context down)."""
    import threading

    def fire():
        sys.stderr.write("bench.py watchdog: no result after %d s, exiting\n" % seconds)
        sys.stderr.flush()
        os._exit(3)
    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()


def main():
    start_watchdog(int(os.environ.get("GR_BENCH_WATCHDOG_S", "900")))
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seq-len", type=int, default=T_FRAMES)
    ap.add_argument("--ref-batch", type=int, default=32, help="sequences per CPU-baseline step (bounded sample)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-ctc", action="store_true")
    ap.add_argument("--ctc-sweep", action="store_true", help="all four label lengths in the config-4 sweep of the 'ctc' object (default: L = 10 and 40)")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only")
    ap.add_argument("--min-warmup", type=int, default=3, help="profiling runs only (ncu launch lists)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
