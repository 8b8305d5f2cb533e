"""FusionTrainer bookkeeping (which towers are launched when, and which prefetch is trusted), on stub objects:
no CUDA.  The numerical equivalence with the serial loop is a GPU test (tests/test_gpu_models.py)."""
import numpy as np


class _StubModel:
    def __init__(self):
        self.log = []

    def sample_regularisers(self, B, T, seed, step, device):
        self.log.append(("reg", step))
        return {"step": step}

    def launch_towers(self, xa, xs, reg, ready=None):
        self.log.append(("towers", reg["step"], id(xa), ready))
        return {"merged": None, "events": [], "inputs": (xa, xs)}

    def join_towers(self, handle):
        self.log.append(("join", id(handle["inputs"][0])))
        return handle["merged"]

    def loss_and_grads(self, xa, xs, labels, il, ll, reg, global_batch=None, towers=None):
        assert towers["inputs"][0] is xa and towers["inputs"][1] is xs      # never someone else's features
        self.log.append(("fusion", reg["step"], id(xa)))
        return np.float32(reg["step"]), ["g%d" % reg["step"]]


class _StubOpt:
    def __init__(self):
        self.seen = []

    def step(self, grads):
        self.seen.append(grads)


class _T:
    """stands in for a (B, T, F) tensor"""
    shape = (4, 10, 3)
    device = "cpu"


def _trainer(hook=None):
    import mgr_b200 as mgr
    m, o = _StubModel(), _StubOpt()
    return mgr.FusionTrainer(m, o, seed=5, global_batch=8, grad_hook=hook), m, o


def test_serial_steps_launch_their_own_towers():
    t, m, o = _trainer()
    a, b = (_T(), _T()), (_T(), _T())
    assert t.step(a + (0, 0, 0)) == 0 and t.step(b + (0, 0, 0)) == 1
    assert [e[:2] for e in m.log] == [("reg", 0), ("towers", 0), ("fusion", 0), ("reg", 1), ("towers", 1), ("fusion", 1)]
    assert o.seen == [["g0"], ["g1"]]


def test_next_batch_towers_are_enqueued_before_this_batch_trains_and_reused():
    t, m, _ = _trainer()
    a, b, c = (_T(), _T()), (_T(), _T()), (_T(), _T())
    t.step(a + (0, 0, 0), next_inputs=b, next_ready="ev1")
    kinds = [e[:2] for e in m.log]
    assert kinds == [("reg", 0), ("towers", 0), ("reg", 1), ("towers", 1), ("fusion", 0)]
    assert m.log[3][3] == "ev1"                       # the copy event reaches launch_towers
    m.log.clear()
    t.step(b + (0, 0, 0), next_inputs=c)
    assert [e[:2] for e in m.log] == [("reg", 2), ("towers", 2), ("fusion", 1)]   # batch b's towers were NOT relaunched
    m.log.clear()
    t.step(c + (0, 0, 0))                             # last batch: nothing to prefetch
    assert [e[:2] for e in m.log] == [("fusion", 2)]


def test_prefetch_for_other_tensors_is_discarded():
    t, m, _ = _trainer()
    a, b, other = (_T(), _T()), (_T(), _T()), (_T(), _T())
    t.step(a + (0, 0, 0), next_inputs=b)
    m.log.clear()
    t.step(other + (0, 0, 0))                         # the caller changed its mind: same step index, other tensors
    # the stale prefetch is JOINED before it is dropped (its buffers are still in use on the side streams)
    assert m.log[0] == ("join", id(b[0]))
    assert [e[:2] for e in m.log[1:]] == [("reg", 1), ("towers", 1), ("fusion", 1)]
    assert m.log[2][2] == id(other[0])


def test_unconsumed_prefetch_is_joined_on_close():
    t, m, _ = _trainer()
    a, b = (_T(), _T()), (_T(), _T())
    t.step(a + (0, 0, 0), next_inputs=b)              # last next_inputs of an epoch, never trained on
    m.log.clear()
    t.close()
    assert m.log == [("join", id(b[0]))]
    t.close()                                         # idempotent
    assert len(m.log) == 1


def test_grad_hook_sits_between_backward_and_optimiser():
    calls = []

    def hook(grads):
        calls.append(list(grads))
        return ["reduced"]
    t, _, o = _trainer(hook)
    t.step((_T(), _T(), 0, 0, 0))
    assert calls == [["g0"]] and o.seen == [["reduced"]]


def test_hook_can_be_ordered_after_the_prefetched_towers():
    order = []
    t, m, _ = _trainer(lambda g: order.append(("hook", len(m.log))) or g)
    t.hook_after_towers = True
    a, b = (_T(), _T()), (_T(), _T())
    t.step(a + (0, 0, 0), next_inputs=b)
    joins = [i for i, e in enumerate(m.log) if e[0] == "join"]
    assert joins and m.log[joins[-1]][1] == id(b[0])          # the calling stream joined batch b's towers ...
    assert order[0][1] == joins[-1] + 1                        # ... right before the hook ran


def test_two_batches_ahead_queue():
    """next_inputs as a LIST: towers of batches n+1 and n+2 are in flight while batch n trains; each is launched once,
    in order, consumed with its own regularisers; with a gradient hook every prefetch is joined before the hook."""
    calls = []
    t, m, o = _trainer(hook=lambda g: (calls.append(len(m.log)), g)[1])
    bs = [(_T(), _T()) for _ in range(5)]
    for n in range(5):
        nxt = [bs[k] for k in (n + 1, n + 2) if k < 5]
        assert t.step(bs[n] + (0, 0, 0), next_inputs=nxt or None, next_ready="ev") == n
    towers = [e[1] for e in m.log if e[0] == "towers"]
    assert towers == [0, 1, 2, 3, 4]                      # every batch's towers launched exactly once, in order
    fus = [e[1] for e in m.log if e[0] == "fusion"]
    assert fus == [0, 1, 2, 3, 4]
    # step 0: towers 0, 1, 2 are enqueued before fusion 0
    first_fusion = m.log.index(("fusion", 0, id(bs[0][0])))
    assert [e[1] for e in m.log[:first_fusion] if e[0] == "towers"] == [0, 1, 2]
    # the hook of step 0 ran after BOTH outstanding prefetches were joined
    joins_before_hook0 = [e for e in m.log[first_fusion:calls[0]] if e[0] == "join"]
    assert [j[1] for j in joins_before_hook0] == [id(bs[1][0]), id(bs[2][0])]
    t.close()


def test_queue_drops_stale_prefetches():
    t, m, _ = _trainer()
    a, b, c, d = [(_T(), _T()) for _ in range(4)]
    t.step(a + (0, 0, 0), next_inputs=[b, c])
    n0 = len(m.log)
    t.step(b + (0, 0, 0), next_inputs=[d])              # c was announced for step 2 but d comes instead
    tail = m.log[n0:]
    assert ("join", id(c[0])) in tail                      # the stale prefetch is joined before it is dropped
    assert [e[1] for e in tail if e[0] == "towers"] == [2] and [e[2] for e in tail if e[0] == "towers"] == [id(d[0])]
    t.step(d + (0, 0, 0))
    assert [e[1] for e in m.log if e[0] == "fusion"] == [0, 1, 2]
