"""Does splitting the batch over CONCURRENT engines help?  Independent clips -> k engines x (B / k) clips on k streams, so
the tail of one chain's kernel (256 tiles over 148 SMs quantise to 86 %) overlaps the other chain's work.
   python tools/two_stream_probe.py [--batch 16] [--ways 2]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import syncfusion_b200 as sf

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--length", type=int, default=262144)
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--calls", type=int, default=4)
ap.add_argument("--ways", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda:0")
cfg = sf.UNetConfig(precision="bf16")
sd = sf.random_state_dict(cfg)
x, ch, e = sf.synthetic_inputs(cfg, a.batch, a.length)
x, e, ch = x.to(dev), e.to(dev), [c.to(dev) for c in ch]


def run(ways):
    ms = []
    for _ in range(ways):
        m = sf.DiffusionModel(cfg, dev); m.load_state_dict(sd); ms.append(m)
    streams = [torch.cuda.Stream(dev) for _ in range(ways)]
    per = a.batch // ways
    parts = [(x[i * per:(i + 1) * per].contiguous(), [c[i * per:(i + 1) * per].contiguous() for c in ch], e[i * per:(i + 1) * per].contiguous()) for i in range(ways)]

    def once():
        outs = []
        for m, s, (xx, cc, ee) in zip(ms, streams, parts):
            with torch.cuda.stream(s):
                outs.append(m.sample(x_noisy=xx, num_steps=a.steps, channels=cc, embedding=ee, embedding_scale=1.0))
        return outs
    for _ in range(2):
        once()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(a.calls):
        outs = once()
    torch.cuda.synchronize()
    dt = (time.time() - t0) / a.calls
    print(f"ways={ways}: {dt*1e3:.1f} ms per {a.batch} clips -> {a.batch/dt:.2f} clips/s", flush=True)
    return torch.cat(outs)

ref = run(1)
for w in sorted({2, a.ways}):
    out = run(w)
    print("   max |diff| vs one engine:", float((out - ref).abs().max()))
