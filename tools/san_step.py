"""compute-sanitizer target: ONE U-Net evaluation with CFG of the full 8-depth architecture (every kernel family: d0 /
rk / sk / attention / prepare / sampler update) at a reduced batch x length, with the grid limited to a few CTAs so that
every persistent CTA walks SEVERAL tiles (ring wrap-around, TMEM double-buffer hand-off, residual-slot recycling are
the code paths a racecheck / synccheck run is for).  Stream launches, no CUDA graph.

    SFB_GRAPH=0 compute-sanitizer --tool synccheck python tools/san_step.py [--batch 2] [--length 32768] [--grid 3]
"""
import argparse
import os
import sys

os.environ.setdefault("SFB_GRAPH", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import syncfusion_b200 as sf

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--length", type=int, default=32768)
ap.add_argument("--grid", type=int, default=3, help="CTA limit of the persistent kernels (0: one per SM)")
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--scale", type=float, default=2.0)
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
dev = torch.device("cuda:0")
cfg = sf.UNetConfig(precision=a.precision)
m = sf.DiffusionModel(cfg, dev)
m.load_state_dict(sf.random_state_dict(cfg))
x, ch, e = sf.synthetic_inputs(cfg, a.batch, a.length)
x, e, ch = x.to(dev), e.to(dev), [c.to(dev) for c in ch]
m.net.debug_set_grid_limit(a.grid)
out = m.sample(x_noisy=x, num_steps=a.steps, channels=ch, embedding=e, embedding_scale=a.scale)
torch.cuda.synchronize()
print("launches", m.net.last_launch_count, "finite", bool(torch.isfinite(out).all()), "out", float(out.abs().mean()))
