"""Run under torchrun (2+ ranks, NCCL): FusionTrainer with the towers one batch ahead and the flat gradient all-reduce
as grad_hook == the same three training steps on ONE rank with the whole batch (rank r owns rows
[r*B/N, (r+1)*B/N), gradients summed before clip/Adam/maxnorm, replicas stay identical).  Regularisers are slices of
one global draw so that both runs see the same masks.  Prints DP_TRAINER_OK on every rank."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import mgr_b200 as mgr
from mgr_b200 import parallel

rank, world, local = parallel.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
GB, T, C, STEPS = 8, 40, 22, 3


def make():
    sp = mgr.UnimodalNet(39, 32, 44, 0.5, (0.4, 0.5, 0.5), seed=1).to(dev)
    sk = mgr.UnimodalNet(20, 16, C, 0.5, (0.6, 0.6, 0.6), seed=2).to(dev)
    return mgr.FusionNet(sp, sk, nb_classes=C, units=12, seed=3).to(dev)


rng = np.random.default_rng(0)
batches = []
for s in range(STEPS):
    xa = torch.tensor(rng.standard_normal((GB, T, 39)).astype(np.float32))
    xs = torch.tensor(rng.standard_normal((GB, T, 20)).astype(np.float32))
    labels = -np.ones((GB, 6), np.float32); ll = np.zeros((GB, 1), np.int64)
    for b in range(GB):
        L = int(rng.integers(1, 7)); labels[b, :L] = rng.integers(0, C - 1, size=L); ll[b, 0] = L
    batches.append((xa, xs, torch.tensor(labels), torch.tensor(np.full((GB, 1), T - 2)), torch.tensor(ll)))


def run(lo, hi, hook_factory):
    model = make()
    full = make()   # only used to draw GLOBAL regularisers with the model's own sampler

    def sliced(B, Tn, seed, step, device):
        reg = full.sample_regularisers(GB, Tn, seed=seed, step=step, device=device)

        def cut(v, batch_axis):
            return v[lo:hi].contiguous() if batch_axis == 0 else v[:, lo:hi].contiguous()
        out = {"sp": {}, "sk": {}}
        for t in ("sp", "sk"):
            for k, v in reg[t].items():
                out[t][k] = cut(v, 0 if k in ("noise", "drop") else 1) if torch.is_tensor(v) else v
        out["m3"] = cut(reg["m3"], 1)
        out["drop"] = cut(reg["drop"], 0)
        out["dropout_masks"] = reg.get("dropout_masks", False)
        return out
    model.sample_regularisers = sliced
    opt = mgr.fusion_optimizer(model)
    trainer = mgr.FusionTrainer(model, opt, seed=11, global_batch=GB, grad_hook=hook_factory(model))
    dev_b = [tuple(t[lo:hi].to(dev) for t in b) for b in batches]
    losses = []
    for s in range(STEPS):
        nxt = (dev_b[s + 1][0], dev_b[s + 1][1]) if s + 1 < STEPS else None
        losses.append(trainer.step(dev_b[s], next_inputs=nxt))
    trainer.close()
    torch.cuda.synchronize()
    return model, torch.cat([l.reshape(-1) for l in losses])


def dp_hook(model):
    bucket = parallel.FlatGradBucket(model.trainable_parameters())

    def hook(grads):
        bucket.pack(grads)
        return bucket.all_reduce()
    return hook


lo, hi = parallel.shard_rows(GB, rank, world)
m_dp, loss_dp = run(lo, hi, dp_hook)
m_one, loss_one = run(0, GB, lambda model: None)          # every rank repeats the single-rank run (no collective in it)
for a, b in zip(m_dp.trainable_parameters(), m_one.trainable_parameters()):
    err = float((a - b).abs().max() / (b.abs().max() + 1e-12))
    assert err < 2e-5, "weights after %d data-parallel steps differ from the single-rank run: %.3e" % (STEPS, err)
assert float((loss_dp - loss_one.reshape(STEPS, GB)[:, lo:hi].reshape(-1)).abs().max()) < 1e-3 * float(loss_one.abs().max())
w = m_dp.blstm_3.kernel.detach().clone(); w0 = w.clone(); dist.broadcast(w0, 0)
assert torch.equal(w, w0), "replicas diverged"
print("DP_TRAINER_OK rank %d of %d rows [%d,%d)" % (rank, world, lo, hi), flush=True)
dist.destroy_process_group()
