"""Device-clock timeline of one sk plan op (CTA 0) on the bench workload.
   python tools/sk_timeline.py --depth 4 --ck conv1 [--batch 16]"""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import syncfusion_b200 as sf

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--length", type=int, default=262144)
ap.add_argument("--ops", default="4:conv1,4:conv2,4:qkv,7:conv2,6:conv1")
a = ap.parse_args()
dev = torch.device("cuda:0")
cfg = sf.UNetConfig(precision="bf16")
m = sf.DiffusionModel(cfg, dev)
m.load_state_dict(sf.random_state_dict(cfg))
x, ch, e = sf.synthetic_inputs(cfg, a.batch, a.length)
x, e, ch = x.to(dev), e.to(dev), [c.to(dev) for c in ch]
net = m.net
t = torch.full((a.batch,), 0.5, device=dev)
net(x, t, embedding=e, embedding_scale=1.0, channels=ch)
ops, ws = net.debug_ops(a.batch, a.length, 0)
names = ["A issue", "B issue", "xform beg/end", "MMA stage", "epi chunk (t0 start,t1 resid,t2 math done,t3 barrier)", "epi store done", "acc wait beg/end", "misc"]
for spec in a.ops.split(","):
    d, ck = spec.split(":")
    idx = next(i for i, o in enumerate(ops) if o["kind"] == "sk" and o["depth"] == int(d) and o["ck"] == ck)
    lib = net._lib
    assert lib.sfb_dbg_sk_timeline(net._h, idx, None, 0) == 0
    net(x, t, embedding=e, embedding_scale=1.0, channels=ch)
    buf = (C.c_longlong * 2048)()
    assert lib.sfb_dbg_sk_timeline(net._h, -1, buf, 2048) == 0
    v = list(buf)
    t0 = v[7 * 256]
    print(f"=== op {idx} depth {d} {ck}: stamps in cycles since prologue end")
    for r in range(8):
        row = [(i, v[r * 256 + i] - t0) for i in range(256) if v[r * 256 + i]]
        print(f"  [{r}] {names[r]}: " + " ".join(f"{i}:{c}" for i, c in row[:80]))
