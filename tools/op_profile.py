"""Per-op device time of one U-Net evaluation on the bench workload (CUDA events around every launch).
   python tools/op_profile.py [--batch 16] [--scale 1.0] > gpurun_out/op_profile.txt"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import syncfusion_b200 as sf

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--length", type=int, default=262144)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
dev = torch.device("cuda:0")
cfg = sf.UNetConfig(precision=a.precision)
m = sf.DiffusionModel(cfg, dev)
m.load_state_dict(sf.random_state_dict(cfg))
x, ch, e = sf.synthetic_inputs(cfg, a.batch, a.length)
x, e, ch = x.to(dev), e.to(dev), [c.to(dev) for c in ch]
for _ in range(2):
    m.sample(x_noisy=x, num_steps=2, channels=ch, embedding=e, embedding_scale=a.scale)
torch.cuda.synchronize()
m.net.profile(True)
m.sample(x_noisy=x, num_steps=2, channels=ch, embedding=e, embedding_scale=a.scale)
torch.cuda.synchronize()
rep = m.net.profile_report()
m.net.profile(False)
ops, _ = m.net.debug_ops(a.batch, a.length, int(a.scale != 1.0))
tot = sum(r["ms"] for r in rep)
print(f"# total {tot:.3f} ms per evaluation, {len(rep)} launches")
print("# idx kind depth stack item ck ms TF/s GB/s share")
agg = {}
for r, o in zip(rep, ops):
    tf = r["flops"] / (r["ms"] * 1e-3) / 1e12 if r["ms"] > 0 else 0
    gb = r["bytes"] / (r["ms"] * 1e-3) / 1e9 if r["ms"] > 0 else 0
    print(f"{r['index']:4d} {r['kind']:10s} d{r['depth']} s{r['stack']} i{r['item']} {o['ck']:8s} {r['ms']*1e3:9.1f} us {tf:8.1f} {gb:8.1f} {100*r['ms']/tot:5.2f}%")
    k = (r["kind"], r["depth"], o["ck"])
    g = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
    g[0] += 1; g[1] += r["ms"]; g[2] += r["flops"]; g[3] += r["bytes"]
print("# --- by (kind, depth, ck)")
for k, g in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[0]:10s} d{k[1]} {k[2]:8s} n={g[0]:3d} {g[1]*1e3:9.1f} us {g[2]/(g[1]*1e-3)/1e12:8.1f} TF/s {g[3]/(g[1]*1e-3)/1e9:8.1f} GB/s {100*g[1]/tot:5.2f}%")
