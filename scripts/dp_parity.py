"""N-rank data-parallel step == 1-rank step on the concatenated batch (SURVEY.md 2.2 parity rule)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import mgr_b200 as mgr
from mgr_b200 import parallel
rank, world, local = parallel.init_from_env()
dev = torch.device("cuda", local); torch.cuda.set_device(local)
GB, T, C = 8, 60, 22
def make():
    sp = mgr.UnimodalNet(39, 32, 44, 0.5, (0.4, 0.5, 0.5), seed=1).to(dev)
    sk = mgr.UnimodalNet(20, 16, C, 0.5, (0.6, 0.6, 0.6), seed=2).to(dev)
    return mgr.FusionNet(sp, sk, nb_classes=C, units=12, seed=3).to(dev)
rng = np.random.default_rng(0)
xa = torch.tensor(rng.standard_normal((GB, T, 39)).astype(np.float32)); xs = torch.tensor(rng.standard_normal((GB, T, 20)).astype(np.float32))
labels = -np.ones((GB, 6), np.float32); ll = np.zeros((GB, 1), np.int64)
for b in range(GB):
    L = int(rng.integers(1, 7)); labels[b, :L] = rng.integers(0, C - 1, size=L); ll[b, 0] = L
il = np.full((GB, 1), T - 2)
m3 = torch.tensor(((rng.random((8, GB, 96)) > 0.5) / 0.5).astype(np.float32))
drop = torch.tensor(((rng.random((GB, T, 24)) > 0.5) / 0.5).astype(np.float32))
def grads_for(lo, hi, model):
    reg = {"sp": {}, "sk": {}, "m3": m3[:, lo:hi].contiguous().to(dev), "drop": drop[lo:hi].contiguous().to(dev)}
    return model.loss_and_grads(xa[lo:hi].to(dev), xs[lo:hi].to(dev), labels[lo:hi], il[lo:hi], ll[lo:hi], reg, global_batch=GB)
model = make()
lo, hi = parallel.shard_rows(GB, rank, world)
loss, g = grads_for(lo, hi, model)
bucket = parallel.FlatGradBucket(model.trainable_parameters()); bucket.pack(g); views = bucket.all_reduce()
ref_model = make()
_, gref = grads_for(0, GB, ref_model)
err = max(float((a - b).abs().max() / (b.abs().max() + 1e-12)) for a, b in zip(views, gref))
print("rank %d world %d rows [%d,%d) max rel grad diff vs single-rank: %.3e" % (rank, world, lo, hi, err))
assert err < 1e-4
opt = mgr.fusion_optimizer(model); opt.step(views)
w = model.blstm_3.kernel.detach().clone(); w0 = w.clone(); dist.broadcast(w0, 0)
assert torch.equal(w, w0), "replicas diverged after the identical update"
print("rank %d: replicas identical after Adam step" % rank)
dist.destroy_process_group()
