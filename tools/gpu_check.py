"""Developer bring-up script (run on the GPU box): staged kernel checks, each printing PASS/FAIL with error
norms.  Each stage is meant to run in its own process (a trapped kernel poisons the CUDA context):
    python tools/gpu_check.py gemm_bf16 | gemm_f32 | attn_bf16 | attn_f32 | unet_bf16 | unet_f32 | sample
"""
from __future__ import annotations

import ctypes as C
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from syncfusion_b200 import _lib

dev = torch.device("cuda:0")


def P(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def gemm_case(lib, bf16, B, L, K1, K2, N, taps, bmod2, bias_mod, gs, use_bias, use_resid, name):
    dt = torch.bfloat16 if bf16 else torch.float32
    g = torch.Generator(device="cpu").manual_seed(1)
    a1 = torch.randn(B, L, K1, generator=g).to(dev).to(dt)
    B2 = bmod2 if K2 else 0
    a2 = torch.randn(max(B2, 1), L, max(K2, 8), generator=g).to(dev).to(dt) if K2 else None
    w = (torch.randn(taps * N, K1 + K2, generator=g) / (taps * (K1 + K2)) ** 0.5).to(dev).to(dt)
    bm = bias_mod if bias_mod else N
    bias = torch.randn(bm, generator=g).to(dev) if use_bias else None
    resid = torch.randn(B, L, N, generator=g).to(dev) if use_resid else None
    out_r = torch.full((B, L, N), float("nan"), device=dev)
    out_t = torch.zeros(B, L, N, device=dev, dtype=dt)
    stats = torch.zeros(B, 8, 2, device=dev, dtype=torch.float64) if gs else None
    rc = lib.sfb_dbg_gemm(int(bf16), P(a1), P(a2), P(w), P(bias), P(resid), P(out_r), P(out_t), P(stats), B, L, K1, K2, N,
                          taps, max(B2, 1), bm, gs, C.c_void_p(0))
    torch.cuda.synchronize()
    # reference (fp32 math on the rounded operands)
    A = a1.float()
    Wf = w.float()
    ref = torch.zeros(B, L, N, device=dev)
    pad = 1 if taps == 3 else 0
    for t in range(taps):
        sh = t - pad
        As = torch.zeros_like(A)
        if sh == 0:
            As = A
        elif sh < 0:
            As[:, 1:] = A[:, :-1]
        else:
            As[:, :-1] = A[:, 1:]
        ref += As @ Wf[t * N:(t + 1) * N, :K1].T
    if K2:
        idx = torch.arange(B, device=dev) % B2
        ref += a2.float()[idx] @ Wf[:N, K1:].T
    if use_bias:
        ref += bias[torch.arange(N, device=dev) % bm]
    if use_resid:
        ref += resid
    e_r, e_t = rel(out_r, ref), rel(out_t.float(), ref)
    ok = rc == 0 and e_r < (2e-3 if not bf16 else 1e-4) and e_t < 1e-2
    msg = f"{name}: rc={rc} rel_r={e_r:.3e} rel_t={e_t:.3e}"
    if gs:
        ch = torch.arange(N, device=dev) % bm
        grp = ch // gs
        s_ref = torch.zeros(B, 8, 2, device=dev, dtype=torch.float64)
        for gi in range(8):
            m = grp == gi
            if m.any():
                s_ref[:, gi, 0] = ref[:, :, m].double().sum(dim=(1, 2))
                s_ref[:, gi, 1] = (ref[:, :, m].double() ** 2).sum(dim=(1, 2))
        e_s = rel(stats, s_ref)
        ok = ok and e_s < 5e-3
        msg += f" rel_stats={e_s:.3e}"
    print(("PASS " if ok else "FAIL ") + msg, flush=True)
    return ok


def stage_gemm(bf16):
    lib = _lib.load()
    cases = [
        # B, L, K1, K2, N, taps, bmod2, bias_mod, gs, bias, resid
        (1, 128, 64, 0, 128, 1, 0, 0, 0, False, False, "plain 128x128x64"),
        (2, 256, 128, 0, 128, 1, 0, 0, 0, True, False, "bias K=128"),
        (2, 256, 256, 0, 256, 1, 0, 0, 0, True, True, "BN=256 resid"),
        (2, 256, 128, 0, 64, 3, 0, 0, 8, True, False, "conv3 N=64 stats"),
        (2, 200, 32, 0, 32, 3, 0, 0, 4, True, True, "conv3 C=32 ragged L (OOB K pad)"),
        (3, 64, 128, 0, 128, 3, 0, 0, 16, True, True, "conv3 L<128"),
        (4, 256, 128, 32, 128, 1, 2, 0, 16, True, True, "inject dual-source bmod"),
        (2, 256, 32, 8, 32, 1, 2, 0, 4, True, True, "inject C=32 ctx=8"),
        (2, 128, 64, 0, 128, 3, 0, 32, 4, True, True, "up-style bias_mod"),
        (2, 512, 1024, 0, 1024, 3, 0, 0, 128, True, True, "conv3 C=1024"),
        (2, 256, 512, 0, 1536, 1, 0, 0, 0, True, False, "qkv N=1536"),
    ]
    ok = True
    for c in cases:
        ok &= gemm_case(lib, bf16, *c)
    return ok


def stage_attn(bf16):
    lib = _lib.load()
    dt = torch.bfloat16 if bf16 else torch.float32
    ok = True
    for (B, N) in [(1, 128), (2, 256), (2, 64), (1, 40), (2, 1024), (1, 2048)]:
        g = torch.Generator().manual_seed(N)
        qkv = (torch.randn(B, N, 1536, generator=g)).to(dev).to(dt)
        qkv[..., :512] *= 2.0
        out = torch.zeros(B, N, 512, device=dev, dtype=dt)
        rc = lib.sfb_dbg_attention(int(bf16), P(qkv), P(out), B, N, C.c_void_p(0))
        torch.cuda.synchronize()
        q, k, v = qkv.float().split(512, dim=-1)
        q, k, v = (t.reshape(B, N, 8, 64).transpose(1, 2) for t in (q, k, v))
        ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, 512)
        e = rel(out.float(), ref)
        good = rc == 0 and e < (1.5e-2 if bf16 else 2e-3)
        if not good:
            o = out.float()
            print(f"   out absmax={o.abs().max().item():.3e} nan={int(torch.isnan(o).sum())} ref absmax={ref.abs().max().item():.3e} "
                  f"out[0,0,:4]={o[0,0,:4].tolist()} ref[0,0,:4]={ref[0,0,:4].tolist()}")
            # V = 1 probe: output must be exactly 1 whatever the softmax does
            q2 = qkv.clone(); q2[..., 1024:] = 1.0
            o2 = torch.zeros_like(out)
            lib.sfb_dbg_attention(int(bf16), P(q2), P(o2), B, N, C.c_void_p(0)); torch.cuda.synchronize()
            print(f"   V=1 probe: min={o2.float().min().item():.4f} max={o2.float().max().item():.4f}")
            # Q = 0 probe: uniform attention -> mean of V
            q3 = qkv.clone(); q3[..., :512] = 0.0
            o3 = torch.zeros_like(out)
            lib.sfb_dbg_attention(int(bf16), P(q3), P(o3), B, N, C.c_void_p(0)); torch.cuda.synchronize()
            vm = q3[..., 1024:].float().mean(dim=1, keepdim=True).expand(B, N, 512)
            print(f"   Q=0 probe: rel to mean(V)={rel(o3.float(), vm):.3e}")
        ok &= good
        print(f"{'PASS' if good else 'FAIL'} attention B={B} N={N}: rc={rc} rel={e:.3e}", flush=True)
    return ok


def stage_unet(precision, cfg_kwargs=None, L=1024, B=2, scale=1.0, upsample_mode="nearest", stress=True):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from tests.util import make_oracle, make_inputs, SMALL
    from tests.trace import trace_unet
    import syncfusion_b200 as sf
    kw = SMALL if cfg_kwargs is None else cfg_kwargs
    om = make_oracle(kw, stress=stress, upsample_mode=upsample_mode)
    x, channels, e = make_inputs(om.net.cfg, B, L)
    cfg = sf.UNetConfig(precision=precision, upsample_mode=upsample_mode, **kw)
    net = sf.UNetV0(cfg, dev)
    net.load_state_dict(om.net.state_dict())
    om = om.to(dev)
    x, e = x.to(dev), e.to(dev)
    channels = [c.to(dev) for c in channels]
    time = torch.full((B,), 0.7, device=dev)
    tr, v_ref = trace_unet(om.net, x, time, e, channels, scale)
    v = net(x, time, embedding=e, embedding_scale=scale, channels=channels)
    torch.cuda.synchronize()
    ev = rel(v, v_ref)
    ed = rel(v - x, v_ref - x)
    print(f"     increment (v - x) rel={ed:.3e}")
    tol = 1e-3 if precision == "fp32" else 3e-2
    ev = max(ev, ed / 10)
    print(f"{'PASS' if ev < tol else 'FAIL'} unet[{precision},{upsample_mode},scale={scale}] v rel={ev:.3e}", flush=True)
    # layer-wise parity: run the plan op by op and compare every op output with the oracle trace
    ops, ws = net.debug_ops(B, L, int(scale != 1.0))
    assert len(ops) == len(tr), (len(ops), len(tr))
    tdt = torch.float32 if precision == "fp32" else torch.bfloat16
    worst = {}
    allok = ev < tol
    verbose = "--all" in sys.argv
    for i, (op, (kinds, ref)) in enumerate(zip(ops, tr)):
        net.debug_set_op_limit(i + 1)
        net(x, time, embedding=e, embedding_scale=scale, channels=channels)
        torch.cuda.synchronize()
        raw = ws[((ws.data_ptr() + 1023) // 1024 * 1024 - ws.data_ptr()) + op["off"]:][: op["nbytes"]]
        got = raw.view(torch.float32 if op["dtype"] == 0 else tdt).reshape(op["rows"], op["cols"]).float()
        err = rel(got, ref.reshape(op["rows"], op["cols"])) if got.numel() == ref.numel() else float("nan")
        bad = not (err < (2e-3 if precision == "fp32" else 5e-2))
        key = (op["kind"], op["depth"])
        worst[key] = max(worst.get(key, 0.0), err if err == err else 9.9)
        if bad or verbose:
            print(f"  [{i:3d}] {'BAD' if bad else 'ok '} {op['kind']:10s} d{op['depth']} s{op['stack']} i{op['item']} "
                  f"[{op['rows']}x{op['cols']}] vs {kinds:16s} rel={err:.3e}", flush=True)
        allok &= not bad
    print("  worst per (kind, depth): " + ", ".join(f"{k[0]}@d{k[1]}={v:.1e}" for k, v in sorted(worst.items())))
    net.debug_set_op_limit(-1)
    return allok



def stage_sample(precision):
    from tests.util import make_oracle, make_inputs, SMALL
    import syncfusion_b200 as sf
    om = make_oracle(SMALL, stress=True)
    B, L, N = 2, 2048, 5
    x, channels, e = make_inputs(om.net.cfg, B, L)
    m = sf.DiffusionModel(sf.UNetConfig(precision=precision, **SMALL), dev)
    m.load_state_dict(om.net.state_dict())
    om = om.to(dev)
    x, e = x.to(dev), e.to(dev)
    channels = [c.to(dev) for c in channels]
    ok = True
    for scale in (1.0, 2.0):
        ref, xs, vs = om.sampler(x, N, channels=channels, embedding=e, embedding_scale=scale, return_trajectory=True)
        out = m.sample(x_noisy=x, num_steps=N, channels=channels, embedding=e, embedding_scale=scale)
        torch.cuda.synchronize()
        err = rel(out, ref)
        good = err < (1e-2 if precision == "fp32" else 5e-2)
        ok &= good
        print(f"{'PASS' if good else 'FAIL'} sample[{precision}] scale={scale} free-running final rel={err:.3e} "
              f"launches={m.net.last_launch_count}", flush=True)
    return ok


if __name__ == "__main__":
    st = sys.argv[1]
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    if st == "gemm_bf16": r = stage_gemm(True)
    elif st == "gemm_f32": r = stage_gemm(False)
    elif st == "attn_bf16": r = stage_attn(True)
    elif st == "attn_f32": r = stage_attn(False)
    elif st == "unet_bf16": r = stage_unet("bf16")
    elif st == "unet_f32": r = stage_unet("fp32")
    elif st == "unet_f32_cfg": r = stage_unet("fp32", scale=2.0)
    elif st == "unet_f32_T": r = stage_unet("fp32", upsample_mode="transpose")
    elif st == "unet_full_bf16": r = stage_unet("bf16", cfg_kwargs={}, L=4096, B=1, scale=2.0)
    elif st == "unet_full_f32": r = stage_unet("fp32", cfg_kwargs={}, L=4096, B=1, scale=2.0)
    elif st == "sample_bf16": r = stage_sample("bf16")
    elif st == "sample_f32": r = stage_sample("fp32")
    else: raise SystemExit(f"unknown stage {st}")
    sys.exit(0 if r else 1)
