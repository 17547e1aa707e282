"""Random-init weights and synthetic inputs of the named architecture (checkpoints and datasets are not available
offline, BASELINE.json).  Host-side product code: no dependency on ``oracle/``.

``random_state_dict`` enumerates every flat C-ABI parameter name (the same set ``sfb_finalize`` requires) with
PyTorch's default initialisers, so the library can be exercised without the oracle; ``synthetic_inputs`` draws
noise / onset-pyramid / embedding tensors with the shapes of exp/train_diffusion_gh.yaml:7-9 and
exp/model/diffusion.yaml:22,32.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
from torch import Tensor

from .config import UNetConfig


def _uniform(g, shape, bound):
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def _conv(g, out_c, in_c, k) -> Tuple[Tensor, Tensor]:
    b = 1.0 / math.sqrt(in_c * k)
    return _uniform(g, (out_c, in_c, k), b), _uniform(g, (out_c,), b)


def _linear(g, out_f, in_f, bias=True):
    b = 1.0 / math.sqrt(in_f)
    return _uniform(g, (out_f, in_f), b), (_uniform(g, (out_f,), b) if bias else None)


def param_shapes(cfg: UNetConfig) -> Dict[str, Tuple[int, ...]]:
    """Flat C-ABI parameter name -> shape for this configuration: exactly the set ``sfb_finalize`` requires (used by
    ``DiffusionModel.load_state_dict(strict=True)`` to report missing / unexpected keys without allocating weights)."""
    sh: Dict[str, Tuple[int, ...]] = {}
    mf, ef = cfg.modulation_features, cfg.embedding_features
    mid = cfg.attention_heads * cfg.attention_features
    sh["time.weights"] = (128,)
    sh["time.linear.weight"], sh["time.linear.bias"] = (mf, 257), (mf,)
    sh["time.mlp.weight"], sh["time.mlp.bias"] = (mf, mf), (mf,)
    sh["fixed_embedding.weight"] = (cfg.embedding_max_length, ef)
    for d in range(cfg.depth):
        c, f, ctx = cfg.channels[d], cfg.factors[d], cfg.context_channels[d]
        cin = cfg.in_channels if d == 0 else cfg.channels[d - 1]
        p = f"d{d}."
        sh[p + "down.weight"], sh[p + "down.bias"] = (c, cin, f), (c,)
        if cfg.upsample_mode == "transpose":
            sh[p + "up.weight"], sh[p + "up.bias"] = (c, cin, f), (cin,)
        else:
            sh[p + "up.conv.weight"], sh[p + "up.conv.bias"] = (cin, c, 3), (cin,)
        sh[p + "skip.weight"], sh[p + "skip.bias"] = (cin, mf), (cin,)
        for stack in ("items_down", "items_up"):
            for i in range(cfg.items[d]):
                q = f"{p}{stack}.{i}."
                for n in ("gn1", "gn2"):
                    sh[q + f"resnet.{n}.weight"] = sh[q + f"resnet.{n}.bias"] = (c,)
                for n in ("conv1", "conv2"):
                    sh[q + f"resnet.{n}.weight"], sh[q + f"resnet.{n}.bias"] = (c, c, 3), (c,)
                sh[q + "mod.linear.weight"], sh[q + "mod.linear.bias"] = (2 * c, mf), (2 * c,)
                if ctx > 0:
                    sh[q + "inject.conv.weight"], sh[q + "inject.conv.bias"] = (c, c + ctx, 1), (c,)
                kinds = []
                if cfg.attentions[d]:
                    kinds.append(("attn", c))
                if cfg.cross_attentions[d]:
                    kinds.append(("xattn", ef))
                for name, cf in kinds:
                    a = q + f"{name}.attn."
                    sh[a + "norm.weight"] = sh[a + "norm.bias"] = (c,)
                    sh[a + "norm_ctx.weight"] = sh[a + "norm_ctx.bias"] = (cf,)
                    sh[a + "to_q.weight"] = (mid, c)
                    sh[a + "to_kv.weight"] = (2 * mid, cf)
                    sh[a + "to_out.weight"] = (c, mid)
    return sh


def random_state_dict(cfg: UNetConfig, seed: int = 0) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    mf, ef = cfg.modulation_features, cfg.embedding_features
    mid = cfg.attention_heads * cfg.attention_features
    sd["time.weights"] = torch.randn(128, generator=g)
    sd["time.linear.weight"], sd["time.linear.bias"] = _linear(g, mf, 257)
    sd["time.mlp.weight"], sd["time.mlp.bias"] = _linear(g, mf, mf)
    sd["fixed_embedding.weight"] = torch.randn(cfg.embedding_max_length, ef, generator=g)
    for d in range(cfg.depth):
        c, f, ctx = cfg.channels[d], cfg.factors[d], cfg.context_channels[d]
        cin = cfg.in_channels if d == 0 else cfg.channels[d - 1]
        p = f"d{d}."
        sd[p + "down.weight"], sd[p + "down.bias"] = _conv(g, c, cin, f)
        if cfg.upsample_mode == "transpose":
            w, _ = _conv(g, c, cin, f)               # ConvTranspose1d weight is [in = C_d, out = C_{d-1}, f]
            sd[p + "up.weight"] = w
            sd[p + "up.bias"] = _uniform(g, (cin,), 1.0 / math.sqrt(c * f))
        else:
            sd[p + "up.conv.weight"], sd[p + "up.conv.bias"] = _conv(g, cin, c, 3)
        sd[p + "skip.weight"], sd[p + "skip.bias"] = _linear(g, cin, mf)
        for stack in ("items_down", "items_up"):
            for i in range(cfg.items[d]):
                q = f"{p}{stack}.{i}."
                for n in ("gn1", "gn2"):
                    sd[q + f"resnet.{n}.weight"] = torch.ones(c)
                    sd[q + f"resnet.{n}.bias"] = torch.zeros(c)
                sd[q + "resnet.conv1.weight"], sd[q + "resnet.conv1.bias"] = _conv(g, c, c, 3)
                sd[q + "resnet.conv2.weight"], sd[q + "resnet.conv2.bias"] = _conv(g, c, c, 3)
                sd[q + "mod.linear.weight"], sd[q + "mod.linear.bias"] = _linear(g, 2 * c, mf)
                if ctx > 0:
                    sd[q + "inject.conv.weight"], sd[q + "inject.conv.bias"] = _conv(g, c, c + ctx, 1)
                kinds = []
                if cfg.attentions[d]:
                    kinds.append(("attn", c))
                if cfg.cross_attentions[d]:
                    kinds.append(("xattn", ef))
                for name, cf in kinds:
                    a = q + f"{name}.attn."
                    sd[a + "norm.weight"], sd[a + "norm.bias"] = torch.ones(c), torch.zeros(c)
                    sd[a + "norm_ctx.weight"], sd[a + "norm_ctx.bias"] = torch.ones(cf), torch.zeros(cf)
                    sd[a + "to_q.weight"], _ = _linear(g, mid, c, bias=False)
                    sd[a + "to_kv.weight"], _ = _linear(g, 2 * mid, cf, bias=False)
                    sd[a + "to_out.weight"], _ = _linear(g, c, mid, bias=False)
    return sd


def synthetic_inputs(cfg: UNetConfig, batch: int, length: int, seed: int = 12345) -> Tuple[Tensor, List[Tensor], Tensor]:
    """noise ~ N(0,1) [B,1,L] (main/generation.py:69); a synthetic onset-encoder pyramid with the shapes of
    ``y_latent['xs'][2:-1]`` (main/generation.py:80) and a unit-norm 512-d embedding [B,1,512] (CLAP outputs are
    unit norm; main/module_diffusion.py:67,71).  Throughput does not depend on the values."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 1, length, generator=g)
    channels = [0.5 * torch.randn(batch, cfg.context_channels[d], cfg.length_at(length, d), generator=g)
                for d in range(cfg.depth)]
    e = torch.randn(batch, 1, cfg.embedding_features, generator=g)
    e = e / e.norm(dim=-1, keepdim=True)
    return x, channels, e
