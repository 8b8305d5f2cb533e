"""BASELINE configs 1 and 2 alone (bench.py's unimodal legs)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
dev = torch.device("cuda:0")
hbm, tc, _ = bench.read_peaks()
for kind in ("skeletal_train", "speech_fwd"):
    r = bench.unimodal_leg(dev, hbm, tc, kind, steps=5, cpu=False)
    print(kind, "%.2f ms/step  %.0f seq/s" % (r["ms_per_step"], r["seq_per_s"]))
    print("   ", r["kernels"])
    print("   ", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r["roofline"].items()})
