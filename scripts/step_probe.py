"""What a rank of an N-GPU run does, on one GPU: the pipelined fusion training step at per-GPU batch B (no
collective), device ms per step against the host time spent enqueueing it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import mgr_b200 as mgr
from mgr_b200 import parallel, _lib
dev = torch.device("cuda:0")
T = int(os.environ.get("T", "1000"))
for B in [int(v) for v in os.environ.get("BS", "32,64,128,256").split(",")]:
    model = mgr.FusionNet().to(dev)
    opt = mgr.fusion_optimizer(model)
    bucket = parallel.FlatGradBucket(model.trainable_parameters())
    xa, xs, lab, il, ll = [t.to(dev) for t in bench.synth_batch(0, B, T)]

    def hook(grads):
        bucket.pack(grads)
        return bucket.all_reduce()      # world size 1: returns the flat views (multi-tensor optimiser epilogue)
    tr = mgr.FusionTrainer(model, opt, seed=1, global_batch=256, grad_hook=hook)
    for _ in range(3):
        tr.step((xa, xs, lab, il, ll), next_inputs=(xa, xs))
    torch.cuda.synchronize()
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.launch_count
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        tr.step((xa, xs, lab, il, ll), next_inputs=(xa, xs))
    e1.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("B=%d T=%d: %.2f ms/step on the device, %.2f ms/step of host enqueue time, %d launches/step -> %.0f seq/s x (256/B) = %.0f"
          % (B, T, ms, t_host * 1e3 / n, (_lib.launch_count - l0) / n, B / (ms * 1e-3), 256 / (ms * 1e-3)), flush=True)
    # phases alone (serial): towers only, fusion only
    torch.cuda.synchronize()
    reg = model.sample_regularisers(B, T, seed=1, step=0, device=dev)
    def towers():
        return model.join_towers(model.launch_towers(xa, xs, reg))
    towers(); torch.cuda.synchronize()
    e0.record(); 
    for _ in range(3): h = towers()
    e1.record(); torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / 3
    hnd = model.launch_towers(xa, xs, reg); model.join_towers(hnd); torch.cuda.synchronize()
    def fusion():
        loss, grads = model.loss_and_grads(xa, xs, lab, il, ll, reg, global_batch=256, towers=hnd)
        opt.step(grads)
    fusion(); torch.cuda.synchronize()
    e0.record()
    for _ in range(3): fusion()
    e1.record(); torch.cuda.synchronize()
    ms_f = e0.elapsed_time(e1) / 3
    print("      towers alone %.2f ms, fusion layer (fwd+loss+bwd+Adam) alone %.2f ms" % (ms_t, ms_f), flush=True)
    tr.close()
    del model, opt, tr
    torch.cuda.empty_cache()
