"""Stand-alone attention kernel at a bench shape: CUDA-event time + (SFB_ATTN_TIMELINE=1) the device-clock stamps of
CTA 0's first softmax warp and MMA thread.   python tools/attn_timeline.py [--batch 16] [--tokens 2048]"""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from syncfusion_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--tokens", type=int, default=2048)
a = ap.parse_args()
lib = _lib.load()
dev = torch.device("cuda:0")
qkv = torch.randn(a.batch, a.tokens, 1536, device=dev).to(torch.bfloat16)
out = torch.zeros(a.batch, a.tokens, 512, device=dev, dtype=torch.bfloat16)
P = lambda t: C.c_void_p(t.data_ptr())
tl = os.environ.pop("SFB_ATTN_TIMELINE", None)
for _ in range(3):
    assert lib.sfb_dbg_attention(1, P(qkv), P(out), a.batch, a.tokens, C.c_void_p(0)) == 0
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    lib.sfb_dbg_attention(1, P(qkv), P(out), a.batch, a.tokens, C.c_void_p(0))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = 4.0 * a.batch * 8 * a.tokens * a.tokens * 64
print(f"attention B={a.batch} N={a.tokens}: {ms*1e3:.1f} us per launch, {fl/ms/1e9:.1f} TF/s")
if tl:
    os.environ["SFB_ATTN_TIMELINE"] = "1"
    lib.sfb_dbg_attention(1, P(qkv), P(out), a.batch, a.tokens, C.c_void_p(0))
    torch.cuda.synchronize()
