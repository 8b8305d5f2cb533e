"""The three reference topologies on the B200 kernels, plus the fusion training step.

  * `SpeechNet`  -- audio_network/speech_lstm_ctc_words.py:32-134: GaussianNoise(0.5) ->
    BLSTM(500, dropout .4) -> BLSTM(500, dropout .5) -> add -> Dropout(.5) -> Dense(44) -> softmax.
  * `SkeletalNet` -- skeletal_network/skeletal_lstm_ctc.py:296-394: GaussianNoise(0.5) ->
    BLSTM(300, .6) -> BLSTM(300, .6) -> add -> Dropout(.6) -> Dense(22) -> softmax.
  * `FusionNet` -- multimodal_fusion/multimodal.py:58-215: the two towers (their BLSTM layers
    re-used and frozen, :109-148) -> concat (speech first, :155-156) -> BLSTM(100, dropout .5) ->
    Dropout(.5) -> Dense(22) -> softmax -> ctc_lambda_func; Adam(1e-4, clipvalue .5, decay 1e-5).
All regularisers are explicit arguments (inject them for parity, or draw them with
`sample_regularisers`, on-device Philox) because K.set_learning_phase(1) keeps them active in
training AND in the frozen towers (speech:40, multimodal.py:66).
"""
import os

import torch
from torch import nn

from . import ops
from .layers import BidirectionalLSTM, DenseSoftmax
from .losses import DROP_FRAMES, KERAS_CTC_EPS, _prep_lengths, _raise_on_status


class UnimodalNet(nn.Module):
    def __init__(self, numfeats, units, nb_classes, noise_std, dropouts, seed=47):
        super().__init__()
        self.numfeats, self.units, self.nb_classes, self.noise_std = numfeats, units, nb_classes, noise_std
        self.p1, self.p2, self.pd = dropouts
        self.blstm_1 = BidirectionalLSTM(numfeats, units, dropout=self.p1, seed=seed, name="blstm_1")
        self.blstm_2 = BidirectionalLSTM(2 * units, units, dropout=self.p2, seed=seed + 7, name="blstm_2")
        self.dense = DenseSoftmax(2 * units, nb_classes, seed=seed)

    @property
    def layers(self):
        """Keras depth-sorted layer list as multimodal.py:109-118 indexes it."""
        return [None, None, self.blstm_1, self.blstm_2]

    def sample_regularisers(self, B, T, seed, step, device, head=True):
        """`head=False`: the tower only (no Dropout mask for the uni-modal Dense head -- the fusion model drops that
        head, multimodal.py:109-118, and drawing the (B,T,2H) mask anyway cost 1.6 GB of writes per step)."""
        reg = {}
        off = int(step) * 64
        if self.noise_std > 0:
            reg["noise"] = ops.gaussian_noise((B, T, self.numfeats), self.noise_std, seed, off, device)
        if self.p1 > 0:
            reg["m1"] = ops.dropout_mask((8, B, self.numfeats), self.p1, seed + 1, off, device)
        if self.p2 > 0:
            reg["m2"] = ops.dropout_mask((8, B, 2 * self.units), self.p2, seed + 2, off, device)
        if self.pd > 0 and head:
            reg["drop"] = ops.dropout_mask((B, T, 2 * self.units), self.pd, seed + 3, off, device)
        reg["dropout_masks"] = True    # m1 / m2 hold only 0 and 1/(1-p): see BidirectionalLSTM.forward
        return reg

    def tower(self, x, reg=None, merged=None, col0=0):
        """GaussianNoise -> BLSTM -> BLSTM -> residual add.  With `merged` (a (B,T,Fo) buffer, inference
        only) the add lands in columns [col0, col0 + 2*units) of it -- the fusion model's concat."""
        reg = reg or {}
        if reg.get("noise") is not None:
            x = ops.add(x, reg["noise"])
        dm = bool(reg.get("dropout_masks", False))
        if (merged is not None and not torch.is_grad_enabled() and os.environ.get("GR_TOWER_AUX", "1") != "0"
                and ops.lstm_aux_supported(x.shape[0], self.units) and merged.is_contiguous() and col0 % 4 == 0):
            # the recurrences write the residual sum themselves: layer 1 stores h into the concat block, layer 2 adds its h
            y1 = self.blstm_1.forward_into(x, reg.get("m1"), dm, merged, col0, accumulate=False)
            self.blstm_2.forward_into(y1, reg.get("m2"), dm, merged, col0, accumulate=True, want_y=False)
            return merged
        y1 = self.blstm_1(x, reg.get("m1"), dm)
        y2 = self.blstm_2(y1, reg.get("m2"), dm)
        if merged is not None:
            return ops.add_into(y1, y2, merged, col0)
        return _add(y1, y2)

    def forward(self, x, reg=None):
        """-> (y_pred softmax probabilities (B,T,C), inner logits)."""
        reg = reg or {}
        res = self.tower(x, reg)
        return self.dense(res, reg.get("drop"))


class _AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return ops.add(a, b)

    @staticmethod
    def backward(ctx, g):
        return g, g


def _add(a, b):
    return _AddFn.apply(a, b)


def SpeechNet(numfeats=39, nb_classes=44, units=500, seed=47):
    return UnimodalNet(numfeats, units, nb_classes, noise_std=0.5, dropouts=(0.4, 0.5, 0.5), seed=seed)


def SkeletalNet(numfeats=20, nb_classes=22, units=300, seed=53):
    return UnimodalNet(numfeats, units, nb_classes, noise_std=0.5, dropouts=(0.6, 0.6, 0.6), seed=seed)


SPLIT_MIN_HALF = 128   # split only batches of more than one 128-row MMA tile: each half still fills its tile


class FusionNet(nn.Module):
    def __init__(self, speech=None, skeletal=None, nb_classes=22, units=100, seed=61):
        super().__init__()
        self.speech = speech if speech is not None else SpeechNet()
        self.skeletal = skeletal if skeletal is not None else SkeletalNet()
        # multimodal.py:135-148: freeze the four tower Bidirectional layers
        for tower in (self.speech, self.skeletal):
            for l in (tower.blstm_1, tower.blstm_2):
                l.trainable = True
                l.forward_layer.trainable = False
                l.backward_layer.trainable = False
            for p in tower.dense.parameters():  # the uni-modal heads are not part of the fusion graph
                p.requires_grad_(False)
        fin = 2 * self.speech.units + 2 * self.skeletal.units
        self.units, self.nb_classes = units, nb_classes
        self._streams = None
        self.blstm_3 = BidirectionalLSTM(fin, units, dropout=0.5, seed=seed, name="blstm_2")
        self.dense = DenseSoftmax(2 * units, nb_classes, seed=seed)

    def trainable_parameters(self):
        return [self.blstm_3.kernel, self.blstm_3.recurrent_kernel, self.blstm_3.bias, self.dense.kernel,
                self.dense.bias]

    def sample_regularisers(self, B, T, seed, step, device):
        off = int(step) * 64
        reg = {"sp": {}, "sk": {}}
        sp = self.speech.sample_regularisers(B, T, seed + 10, step, device, head=False)
        sk = self.skeletal.sample_regularisers(B, T, seed + 20, step, device, head=False)
        sk.pop("noise", None)  # GaussianNoise(0.0) on the skeletal branch (multimodal.py:105)
        reg["sp"], reg["sk"] = sp, sk
        reg["m3"] = ops.dropout_mask((8, B, self.blstm_3.input_dim), self.blstm_3.dropout, seed + 30, off, device)
        reg["dropout_masks"] = True
        reg["drop"] = ops.dropout_mask((B, T, 2 * self.units), 0.5, seed + 31, off, device)
        return reg

    def alloc_merged(self, xa):
        """The Merge(concat) buffer of a batch (allocated on the calling stream)."""
        return torch.empty(xa.shape[:2] + (2 * self.speech.units + 2 * self.skeletal.units,), dtype=torch.float32,
                           device=xa.device)

    def launch_towers(self, xa, xs, reg=None, ready=None, after=None, merged=None):
        """Enqueue the two frozen towers (multimodal.py:109-118) and the Merge(concat) (`:155`) WITHOUT joining
        the calling stream: returns a handle for `join_towers`.  The towers are independent until the concat, so
        they run on side streams; because they are frozen (`:135-148`) their output for batch n+1 does not depend
        on the weight update of step n either, which is what `FusionTrainer` uses to run them one batch ahead.
        `ready`: CUDA event after which xa/xs are valid (inputs copied on another stream).
        `after` + `merged`: the side streams wait for the event `after` (recorded on the calling stream when `reg` and
        the pre-allocated `merged` buffer were ready) instead of for everything enqueued on the calling stream so far --
        lets the caller enqueue other work first without delaying the towers on the GPU (FusionTrainer)."""
        reg = reg or {}
        with torch.no_grad():
            fa, fs = 2 * self.speech.units, 2 * self.skeletal.units
            fused = fa % 4 == 0 and fs % 4 == 0   # 16-byte column blocks: the adds write the concat directly
            if not fused or os.environ.get("GR_TOWER_STREAMS", "1") == "0":
                if ready is not None:
                    torch.cuda.current_stream().wait_event(ready)
                merged = torch.empty(xa.shape[:2] + (fa + fs,), dtype=torch.float32, device=xa.device) if fused else None
                ra = self.speech.tower(xa, reg.get("sp"), merged, 0)
                rs = self.skeletal.tower(xs, reg.get("sk"), merged, fa)
                return {"merged": merged if fused else ops.concat2(ra, rs), "events": [], "inputs": (xa, xs)}
            if merged is None:
                merged = torch.empty(xa.shape[:2] + (fa + fs,), dtype=torch.float32, device=xa.device)
                after = None
            cur = torch.cuda.current_stream()
            if self._streams is None:
                # (higher stream priority for the towers was measured: no gain)
                prio = int(os.environ.get("GR_TOWER_PRIO", "0"))
                self._streams = tuple(torch.cuda.Stream(priority=prio) for _ in range(4))
            sa, sb, sc, sd = self._streams
            B = xa.shape[0]
            work = []
            # lstm_tcu.cu (default) runs a whole 256-sequence layer on 64 / 38 CTAs, so both towers fit side by side
            # unsplit (43.0 ms/step against 48.9 with halves); the half-batch schedule below is what the round-1
            # kernel (GR_LSTM_TCU=0: 128 CTAs per speech layer) needs.
            split = os.environ.get("GR_TOWER_SPLIT", "2" if os.environ.get("GR_LSTM_TCU") == "0" else "0")
            if B >= 2 * SPLIT_MIN_HALF and split in ("1", "2"):
                # The speech recurrence of a 256-sequence batch holds 128 SMs (2 directions x 2 batch tiles x 32
                # unit slices, one CTA per SM), so the skeletal recurrence (76 CTAs) cannot run beside it.  Two
                # half-batch speech towers (64 CTAs each) can each share the GPU with it: 55.0 -> 52.8 ms/step.
                h = B // 2
                sp = reg.get("sp") or {}

                def _cut(d, lo, hi):
                    return {k: (v if not torch.is_tensor(v) else v[lo:hi] if k == "noise" else v[:, lo:hi].contiguous())
                            for k, v in d.items()}

                def half(lo, hi):
                    return _cut(sp, lo, hi)
                sk = reg.get("sk") or {}

                def half_sk(lo, hi):
                    return _cut(sk, lo, hi)
                if split == "2":   # the skeletal tower in halves as well (38 CTAs each): 51.3 -> 49.2 ms/step
                    work = [(sa, lambda: self.speech.tower(xa[:h], half(0, h), merged[:h], 0)),
                            (sb, lambda: self.skeletal.tower(xs[:h], half_sk(0, h), merged[:h], fa)),
                            (sc, lambda: self.speech.tower(xa[h:], half(h, B), merged[h:], 0)),
                            (sd, lambda: self.skeletal.tower(xs[h:], half_sk(h, B), merged[h:], fa))]
                else:
                    work = [(sa, lambda: self.speech.tower(xa[:h], half(0, h), merged[:h], 0)),
                            (sb, lambda: self.skeletal.tower(xs, reg.get("sk"), merged, fa)),
                            (sc, lambda: self.speech.tower(xa[h:], half(h, B), merged[h:], 0))]
            else:
                work = [(sa, lambda: self.speech.tower(xa, reg.get("sp"), merged, 0)),
                        (sb, lambda: self.skeletal.tower(xs, reg.get("sk"), merged, fa))]
            events = []
            for s_, fn in work:
                if after is not None:
                    s_.wait_event(after)
                else:
                    s_.wait_stream(cur)
                if ready is not None:
                    s_.wait_event(ready)
                    xa.record_stream(s_)
                    xs.record_stream(s_)
                with torch.cuda.stream(s_):
                    fn()
                    events.append(s_.record_event())
            return {"merged": merged, "events": events, "inputs": (xa, xs)}

    @staticmethod
    def join_towers(handle):
        cur = torch.cuda.current_stream()
        for ev in handle["events"]:
            cur.wait_event(ev)
        return handle["merged"]

    def merged(self, xa, xs, reg=None):
        return self.join_towers(self.launch_towers(xa, xs, reg))

    def forward(self, xa, xs, reg=None):
        reg = reg or {}
        m = self.merged(xa, xs, reg)
        y3 = self.blstm_3(m, reg.get("m3"), bool(reg.get("dropout_masks", False)))
        return self.dense(y3, reg.get("drop"))

    # ------------------------------------------------------------------ explicit training step
    def loss_and_grads(self, xa, xs, labels, input_length, label_length, reg=None, global_batch=None,
                       check=False, eps=KERAS_CTC_EPS, towers=None):
        """One forward+backward of the Keras objective mean_b ctc_loss_b (dummy loss, speech:131).
        Returns (loss (B,), [grads in trainable_parameters() order]).  `global_batch` = divisor of
        the mean (the data-parallel global batch; default = local B).  `towers`: a `launch_towers` handle for
        these inputs (the frozen towers already enqueued, e.g. one batch ahead)."""
        reg = reg or {}
        B, T = xa.shape[0], xa.shape[1]
        gb = float(global_batch or B)
        lab_i, il, ll = _prep_lengths(labels, input_length, label_length, xa.device)
        m = self.join_towers(towers) if towers is not None else self.merged(xa, xs, reg)
        l3 = self.blstm_3
        y3 = l3(m, reg.get("m3"), bool(reg.get("dropout_masks", False)))   # autograd node (kernel-backed)
        H2 = 2 * self.units
        y3_2d = y3.detach().reshape(B * T, H2)
        dm = reg.get("drop")
        dm2 = None if dm is None else dm.reshape(B * T, H2)
        logits, _ = ops.dense_softmax_fwd(y3_2d, self.dense.kernel.detach(), self.dense.bias.detach(), dm2,
                                          want_probs=False)
        upstream = torch.full((B,), 1.0 / gb, dtype=torch.float32, device=xa.device)
        loss, g_logits, status = ops.ctc_loss_grad(logits.reshape(B, T, self.nb_classes), lab_i, ll, il, True,
                                                   drop_frames=DROP_FRAMES, eps=eps, upstream=upstream)
        if check:
            _raise_on_status(status)
        dWd, dbd, dy3 = ops.dense_bwd(y3_2d, self.dense.kernel.detach(), g_logits.reshape(B * T, -1), dm2,
                                      want_dx=True)
        gW, gU, gb_ = torch.autograd.grad(y3, [l3.kernel, l3.recurrent_kernel, l3.bias],
                                          grad_outputs=dy3.reshape(B, T, H2))
        return loss, [gW, gU, gb_, dWd, dbd]


class KerasAdam:
    """Adam(lr, clipvalue, decay) + maxnorm(3) on LSTM `kernel`s, as compiled at multimodal.py:206-213."""

    def __init__(self, params, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-7, decay=0.0, clipvalue=0.0,
                 maxnorm_params=(), max_norm=3.0):
        self.params = list(params)
        self.lr, self.beta1, self.beta2, self.eps, self.decay, self.clipvalue = lr, beta1, beta2, eps, decay, clipvalue
        self.maxnorm_ids = {id(p) for p in maxnorm_params}
        self.max_norm = max_norm
        # moments live in two flat buffers (views per parameter), so that a flat gradient bucket can be applied in one launch
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.m_flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v_flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m, self.v, o = [], [], 0
        for p in self.params:
            self.m.append(self.m_flat[o:o + p.numel()].view(p.shape))
            self.v.append(self.v_flat[o:o + p.numel()].view(p.shape))
            o += p.numel()
        self.iterations = 0

    def step(self, grads):
        """`grads`: one tensor per parameter.  If they are the views of a flat bucket (`parallel.FlatViews`, what the
        data-parallel hook returns after the all-reduce) and the parameters are dense, the whole update -- clipvalue +
        Adam for every tensor -- is ONE kernel launch, followed by the maxnorm constraint of the LSTM kernels."""
        flat = getattr(grads, "flat", None)
        if (flat is not None and flat.is_cuda and flat.numel() == self.m_flat.numel() and len(self.params) <= 8
                and all(p.data.is_contiguous() for p in self.params)):
            ops.adam_flat([p.data for p in self.params], flat, self.m_flat, self.v_flat, self.iterations, self.lr,
                          self.beta1, self.beta2, self.eps, self.decay, self.clipvalue)
            for p in self.params:
                if id(p) in self.maxnorm_ids:
                    ops.maxnorm(p.data, self.max_norm)
        else:
            for p, g, m, v in zip(self.params, grads, self.m, self.v):
                mn = self.max_norm if id(p) in self.maxnorm_ids else 0.0
                ops.adam_step(p.data, g, m, v, self.iterations, self.lr, self.beta1, self.beta2, self.eps, self.decay,
                              self.clipvalue, mn)
        self.iterations += 1


class FusionTrainer:
    """Training loop body of multimodal.py (`fit_generator` step: forward, CTC objective, Adam/clip/maxnorm) with
    the frozen towers pipelined ONE BATCH AHEAD: while the fusion BLSTM of batch n trains on the calling stream,
    the towers of batch n+1 (whose weights never change, multimodal.py:135-148) already run on the side streams.
    Regularisers are sampled per step index exactly as without the pipeline, so losses and updates are identical.

        trainer = FusionTrainer(model, opt, seed=..., global_batch=...)
        loss = trainer.step(batch_n, next_inputs=(xa_next, xs_next))     # batch = (xa, xs, labels, il, ll)
    """

    def __init__(self, model, opt, seed=0, global_batch=None, grad_hook=None, hook_after_towers=True):
        self.model, self.opt, self.seed, self.global_batch = model, opt, seed, global_batch
        self.grad_hook = grad_hook        # e.g. pack + all-reduce of the flat gradient bucket (data parallel)
        # True (default): the calling stream waits for the prefetched towers before the hook runs, so that a
        # collective in the hook never shares the GPU with the (cooperatively launched, spinning) recurrence
        # kernels -- DESIGN 5.  The next fusion step needs those towers anyway, so the wait is off the critical path.
        self.hook_after_towers = hook_after_towers
        self.step_no = 0                  # next step to be trained
        self._queue = []                  # prefetched towers, oldest first: (step index, reg, towers handle)
        # True: the towers of the NEXT batches are enqueued AFTER this batch's fusion layer (their regularisers and concat
        # buffer are prepared before it and the side streams wait only for that point), so a host that synchronises on
        # every step's loss does not leave the GPU idle while it enqueues ~100 tower launches first (e2e +2.5 %)
        self.late_towers = os.environ.get("GR_TOWERS_LATE", "1") != "0"

    @property
    def _pending(self):                   # (the next step's prefetch, if any; kept for introspection / tests)
        return self._queue[0] if self._queue else None

    def _launch(self, xa, xs, step, ready=None, defer=False):
        B, T = xa.shape[0], xa.shape[1]
        reg = self.model.sample_regularisers(B, T, seed=self.seed, step=step, device=xa.device)
        if defer and getattr(xa, "is_cuda", False) and hasattr(self.model, "alloc_merged"):
            merged = self.model.alloc_merged(xa)
            after = torch.cuda.current_stream().record_event()
            return step, reg, {"merged": None, "events": [], "inputs": (xa, xs), "deferred": (merged, after, ready)}
        return step, reg, self.model.launch_towers(xa, xs, reg, ready)

    def _fire_deferred(self):
        for i, (step, reg, h) in enumerate(self._queue):
            if "deferred" in h:
                merged, after, ready = h["deferred"]
                xa, xs = h["inputs"]
                self._queue[i] = (step, reg, self.model.launch_towers(xa, xs, reg, ready, after=after, merged=merged))

    @staticmethod
    def _matches(entry, step, xa, xs):
        ins = entry[2].get("inputs", (None, None))
        return entry[0] == step and ins[0] is xa and ins[1] is xs

    def step(self, batch, next_inputs=None, next_ready=None):
        """Train on `batch`; `next_inputs` = (xa, xs) of the following batch (its towers are enqueued now), or a LIST of
        the following batches in order (towers two or more batches ahead: more independent work to fill the SMs with;
        single-GPU only -- with a gradient hook every prefetched tower is joined before the collective);
        `next_ready` = CUDA event after which the LAST of those tensors is valid (e.g. recorded on a copy stream), or a
        list with one event (or None) per announced batch."""
        xa, xs, labels, il, ll = batch
        if next_inputs is None:
            upcoming = []
        elif not isinstance(next_inputs[0], (tuple, list)):
            upcoming = [next_inputs]
        else:
            upcoming = list(next_inputs)
        q = self._queue
        if not q or not self._matches(q[0], self.step_no, xa, xs):
            # cold start, or the prefetch was for other tensors.  A stale prefetch is JOINED before it is dropped:
            # its `merged` / regulariser buffers were allocated on this stream while the side streams still work on
            # them; releasing them un-joined would hand the blocks back to this stream's allocator pool mid-flight.
            self._drain_all()
            q = self._queue = [self._launch(xa, xs, self.step_no)]
        # keep what is still valid of the prefetches beyond the head, (re)launch the rest
        for i, (nxa, nxs) in enumerate(upcoming):
            idx = self.step_no + 1 + i
            if len(q) > i + 1 and self._matches(q[i + 1], idx, nxa, nxs):
                continue
            for stale in q[i + 1:]:
                self._drain(stale)
            del q[i + 1:]
            if isinstance(next_ready, (tuple, list)):
                ready = next_ready[i]
            else:
                ready = next_ready if i == len(upcoming) - 1 else None
            q.append(self._launch(nxa, nxs, idx, ready, defer=self.late_towers))
        if len(q) > len(upcoming) + 1:      # prefetches for batches that are no longer announced
            for stale in q[len(upcoming) + 1:]:
                self._drain(stale)
            del q[len(upcoming) + 1:]
        _, reg, towers = q.pop(0)
        loss, grads = self.model.loss_and_grads(xa, xs, labels, il, ll, reg, global_batch=self.global_batch,
                                                towers=towers)
        self._fire_deferred()
        if self.grad_hook is not None:
            if self.hook_after_towers:
                for pend in q:
                    self.model.join_towers(pend[2])
            grads = self.grad_hook(grads)
        self.opt.step(grads)
        self.step_no += 1
        return loss

    def _drain(self, pend):
        if pend is not None:
            self.model.join_towers(pend[2])

    def _drain_all(self):
        q, self._queue = self._queue, []
        for pend in q:
            self._drain(pend)

    def close(self):
        """Join prefetches that will never be consumed (last `next_inputs` of an epoch) before their buffers go."""
        self._drain_all()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fusion_optimizer(model):
    """Adam(lr=0.0001, clipvalue=0.5, decay=1e-5) over the fusion BLSTM + Dense (multimodal.py:206-208)."""
    return KerasAdam(model.trainable_parameters(), lr=1e-4, clipvalue=0.5, decay=1e-5,
                     maxnorm_params=[model.blstm_3.kernel], max_norm=3.0)
