"""Re-runs the oracle U-Net op by op and records a named checkpoint after every sub-step (down, gn1, conv1, gn2,
conv2, mod, inject, attn_ln, qkv, attn, out, up), so every plan op of ``libsyncfusion_b200.so`` can be compared with
the oracle tensor it must reproduce (layer-wise parity).  Fused plan ops name the checkpoint their
output equals; the checkpoints a fused op never materialises are skipped.  Pure oracle math; ``test_oracle.py`` checks that the trace's final tensor equals
``UNetV0.forward``."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn.functional as F


def _nlc(t):                       # [B, C, L] -> [B*L, C]
    return t.transpose(1, 2).reshape(-1, t.shape[1])


def _trace_item(it, x, features, embedding, channels, out: List[Tuple[str, torch.Tensor]]):
    r = it.resnet
    t1 = F.silu(r.gn1(x)); out.append(("gn1", _nlc(t1)))
    t2 = r.conv1(t1); out.append(("conv1", _nlc(t2)))
    t3 = F.silu(r.gn2(t2)); out.append(("gn2", _nlc(t3)))
    h = r.conv2(t3) + x; out.append(("conv2", _nlc(h)))
    m = it.mod(h, features); out.append(("mod", _nlc(m)))
    i = it.inject(m, channels)
    if it.attn is None:
        if it.xattn is not None:
            i = it.xattn(i, embedding)
        out.append(("inject", _nlc(i)))
        return i
    out.append(("inject", _nlc(i)))
    a = it.attn.attn
    xt = i.transpose(1, 2)
    out.append(("attn_ln", F.layer_norm(xt, (xt.shape[-1],)).reshape(-1, xt.shape[-1])))
    q = a.to_q(a.norm(xt)); kv = a.to_kv(a.norm_ctx(xt))
    out.append(("qkv", torch.cat([q, kv], dim=-1).reshape(-1, q.shape[-1] + kv.shape[-1])))
    k, v = kv.chunk(2, dim=-1)
    b, n, _ = q.shape
    hq, hk, hv = (t.reshape(b, n, a.num_heads, -1).transpose(1, 2) for t in (q, k, v))
    o = F.scaled_dot_product_attention(hq, hk, hv).transpose(1, 2).reshape(b * n, -1)
    out.append(("attn", o))
    y = it.attn(i)
    if it.xattn is not None:
        y = it.xattn(y, embedding)
    out.append(("out", _nlc(y)))
    return y


def _trace_block(blk, x, features, embedding, channels, out, depth=0):
    y = blk.down(x); out.append(("down", _nlc(y)))
    for it in blk.items_down:
        y = _trace_item(it, y, features, embedding, channels, out)
    if blk.inner is not None:
        y = _trace_block(blk.inner, y, features, embedding, channels, out, depth + 1)
    for it in blk.items_up:
        y = _trace_item(it, y, features, embedding, channels, out)
    y = blk.up(y)
    s = blk.skip(F.silu(features))[:, :, None]
    r = x + s * y
    out.append(("up0", r.reshape(r.shape[0], -1)) if depth == 0 else ("up", _nlc(r)))
    return r


@torch.no_grad()
def trace_unet(net, x, time, embedding, channels, embedding_scale=1.0):
    """Returns (list of (kinds, [rows, cols] tensor) in plan order, v).  With CFG the rows are [cond; uncond]."""
    features = net.time(time)
    if embedding_scale == 1.0:
        out: List[Tuple[str, torch.Tensor]] = []
        v = _trace_block(net.blocks, x, features, embedding, channels, out)
        return out, v
    b, m = embedding.shape[0], embedding.shape[1]
    mask = net.fixed_embedding(torch.arange(m, device=x.device))[None].expand(b, -1, -1)
    o1, o2 = [], []
    v1 = _trace_block(net.blocks, x, features, embedding, channels, o1)
    v2 = _trace_block(net.blocks, x, features, mask, channels, o2)
    out = []
    for (k, a), (_, c) in zip(o1, o2):
        # rows are clip-major inside each branch: [B*L, C] cond then [B*L, C] uncond
        out.append((k, torch.cat([a, c], dim=0)))
    return out, v2 + (v1 - v2) * embedding_scale
