"""Error of the general cross-attention path vs the oracle as a function of M, init and precision (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import syncfusion_b200 as sf
from tests.util import SMALL, make_inputs, make_oracle, rel_l2
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
for stress in (False, True):
    for precision in ("fp32", "bf16"):
        cfgk = dict(SMALL, embedding_max_length=8)
        om = make_oracle(cfgk, stress=stress)
        m = sf.DiffusionModel(sf.UNetConfig(precision=precision, **cfgk), dev)
        m.load_state_dict(om.net.state_dict())
        om = om.to(dev)
        om64 = make_oracle(cfgk, stress=stress).double().to(dev)
        for M in (1, 2, 4, 8):
            B, L = 2, 1024
            x, ch, _ = make_inputs(om.net.cfg, B, L)
            x, ch = x.to(dev), [c.to(dev) for c in ch]
            g = torch.Generator().manual_seed(M)
            e = torch.randn(B, M, 512, generator=g); e = (e / e.norm(dim=-1, keepdim=True)).to(dev)
            t = torch.tensor([0.7, 0.3], device=dev)
            with torch.no_grad():
                v_ref = om.net(x, t, embedding=e, embedding_scale=1.0, channels=ch)
                v64 = om64.net(x.double(), t.double(), embedding=e.double(), embedding_scale=1.0, channels=[c.double() for c in ch])
            v = m.net(x, t, embedding=e, embedding_scale=1.0, channels=ch)
            print(f"stress={stress} {precision} M={M}: ours vs fp32 oracle {rel_l2(v, v_ref):.2e}  ours vs fp64 oracle {rel_l2(v, v64):.2e}  fp32 oracle vs fp64 {rel_l2(v_ref, v64):.2e}", flush=True)
