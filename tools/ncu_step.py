"""Short profiling target: a few sampler steps of the bench workload (B clips, L = 262144), for ncu launch lists and
`ncu --set full` captures.   python tools/ncu_step.py [--batch 16] [--steps 2] [--scale 1.0] [--precision bf16]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import syncfusion_b200 as sf

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--length", type=int, default=262144)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
dev = torch.device("cuda:0")
cfg = sf.UNetConfig(precision=a.precision)
m = sf.DiffusionModel(cfg, dev)
m.load_state_dict(sf.random_state_dict(cfg))
x, ch, e = sf.synthetic_inputs(cfg, a.batch, a.length)
x, e, ch = x.to(dev), e.to(dev), [c.to(dev) for c in ch]
out = m.sample(x_noisy=x, num_steps=a.steps, channels=ch, embedding=e, embedding_scale=a.scale)
torch.cuda.synchronize()
print("launches", m.net.last_launch_count, "out", float(out.abs().mean()))
