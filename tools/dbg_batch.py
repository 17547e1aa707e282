import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import syncfusion_b200 as sf
from tests.util import SMALL, make_inputs, make_oracle, rel_l2
dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
om = make_oracle(SMALL, stress=True)
m = sf.DiffusionModel(sf.UNetConfig(precision=prec, **SMALL), dev)
m.load_state_dict(om.net.state_dict())
x, ch, e = make_inputs(om.net.cfg, 3, 1024)
x, e, ch = x.to(dev), e.to(dev), [c.to(dev) for c in ch]
net = m.net
t3 = torch.full((3,), 0.7, device=dev)
scale = 2.0
def grab(B, xs, ts, es, chs):
    ops, ws = net.debug_ops(B, 1024, 1)
    pad = (ws.data_ptr() + 1023) // 1024 * 1024 - ws.data_ptr()
    outs = []
    tdt = torch.float32 if prec == "fp32" else torch.bfloat16
    for i, op in enumerate(ops):
        net.debug_set_op_limit(i + 1)
        net(xs, ts, embedding=es, embedding_scale=scale, channels=chs)
        torch.cuda.synchronize()
        raw = ws[pad + op["off"]: pad + op["off"] + op["nbytes"]]
        outs.append(raw.view(torch.float32 if op["dtype"] == 0 else tdt).reshape(op["rows"], op["cols"]).float().clone())
    net.debug_set_op_limit(-1)
    return ops, outs
ops3, o3 = grab(3, x, t3, e, ch)
ops3b, o3b = grab(3, x, t3, e, ch)
ops1, o1 = grab(1, x[1:2], t3[1:2], e[1:2], [c[1:2] for c in ch])
for i, (op, a, a2, b) in enumerate(zip(ops3, o3, o3b, o1)):
    if op["kind"] == "d0_up":
        sel = a[[1, 4]]; sel2 = a2[[1, 4]]
    else:
        rows = a.shape[0] // 6
        sel = torch.cat([a[rows:2 * rows], a[4 * rows:5 * rows]]); sel2 = torch.cat([a2[rows:2 * rows], a2[4 * rows:5 * rows]])
    d = rel_l2(sel, b); d2 = rel_l2(a2, a)
    if d > 1e-6 or d2 > 1e-6 or "--all" in sys.argv:
        print(f"[{i:3d}] {op['kind']:9s} d{op['depth']} s{op['stack']} i{op['item']} batch3-vs-single={d:.3e} rerun={d2:.3e}")
print("done")
