"""Restatement of ``a_unet`` / ``audio_diffusion_pytorch.UNetV0`` (oracle; test infrastructure only).

Parity unpinned: the upstream packages are absent from /root/reference (see
``oracle/__init__.py``).  Every class cites the call site / config line of the
reference that it serves and the SURVEY.md Appendix-A paragraph it restates.

Layout follows the reference: activations are NCL (``"b c t"``,
``main/module_diffusion.py:107``); attention and modulation items work on
``"b t c"`` internally (upstream ``Packed``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor


@dataclass(frozen=True)
class UNetConfig:
    """Field-for-field mirror of ``exp/model/diffusion.yaml:15-33`` (+ upstream defaults)."""

    in_channels: int = 1
    channels: Sequence[int] = (8, 32, 64, 128, 256, 512, 1024, 1024)
    factors: Sequence[int] = (1, 4, 4, 4, 2, 2, 2, 2)
    items: Sequence[int] = (1, 2, 2, 2, 2, 2, 2, 4)
    attentions: Sequence[int] = (0, 0, 0, 0, 1, 1, 1, 1)
    cross_attentions: Sequence[int] = (1, 1, 1, 1, 1, 1, 1, 1)
    context_channels: Sequence[int] = (2, 8, 16, 32, 64, 128, 256, 256)
    attention_heads: int = 8
    attention_features: int = 64
    embedding_features: int = 512
    embedding_max_length: int = 1
    use_embedding_cfg: bool = True
    resnet_groups: int = 8                # upstream default (A.7)
    modulation_features: int = 1024       # upstream default (A.3)
    upsample_mode: str = "nearest"        # A.6: "nearest" (interpolate + conv3) or "transpose"

    @property
    def depth(self) -> int:
        return len(self.channels)

    def length_at(self, length: int, d: int) -> int:
        f = 1
        for i in range(d + 1):
            f *= self.factors[i]
        return length // f

    @property
    def total_factor(self) -> int:
        f = 1
        for x in self.factors:
            f *= x
        return f


# --------------------------------------------------------------------------- items (A.7)


class ResnetItem(nn.Module):
    """A.7 ``ResnetItem``: x + Conv3(SiLU(GN8(Conv3(SiLU(GN8(x)))))); identity shortcut (in == out)."""

    def __init__(self, channels: int, groups: int):
        super().__init__()
        self.gn1 = nn.GroupNorm(groups, channels)
        self.conv1 = nn.Conv1d(channels, channels, 3, padding=1)
        self.gn2 = nn.GroupNorm(groups, channels)
        self.conv2 = nn.Conv1d(channels, channels, 3, padding=1)

    def forward(self, x: Tensor) -> Tensor:
        h = self.conv1(F.silu(self.gn1(x)))
        h = self.conv2(F.silu(self.gn2(h)))
        return h + x


class ModulationItem(nn.Module):
    """A.7 ``ModulationItem``: LN_C(x, no affine) * (1 + scale) + shift, (scale, shift) = Linear(SiLU(features))."""

    def __init__(self, channels: int, features: int):
        super().__init__()
        self.channels = channels
        self.linear = nn.Linear(features, 2 * channels)

    def forward(self, x: Tensor, features: Tensor) -> Tensor:
        xt = x.transpose(1, 2)                                   # Packed: "b c t -> b t c"
        scale, shift = self.linear(F.silu(features))[:, None, :].chunk(2, dim=-1)
        y = F.layer_norm(xt, (self.channels,)) * (1 + scale) + shift
        return y.transpose(1, 2)


class InjectChannelsItem(nn.Module):
    """A.7 ``InjectChannelsItem``: Conv1x1(cat[x, channels[depth]]) + x, with upstream's shape asserts."""

    def __init__(self, channels: int, depth: int, context_channels: int):
        super().__init__()
        self.depth = depth
        self.context_channels = context_channels
        self.conv = nn.Conv1d(channels + context_channels, channels, 1)

    def forward(self, x: Tensor, channels: Sequence[Tensor]) -> Tensor:
        assert channels is not None and self.depth < len(channels), "context `channels` missing for this depth"
        ctx = channels[self.depth]
        assert ctx.shape == (x.shape[0], self.context_channels, x.shape[2]), (
            f"channels[{self.depth}] must be {(x.shape[0], self.context_channels, x.shape[2])}, got {tuple(ctx.shape)}")
        return self.conv(torch.cat([x, ctx], dim=1)) + x


class Attention(nn.Module):
    """A.7 ``Attention``: skip + W_o softmax(q k^T / sqrt(64)) v; separate affine LayerNorms on x and context;
    all three projections bias-free; heads "b n (h d) -> b h n d"."""

    def __init__(self, features: int, head_features: int, num_heads: int, context_features: Optional[int] = None):
        super().__init__()
        mid = head_features * num_heads
        ctx = context_features if context_features is not None else features
        self.num_heads = num_heads
        self.scale = head_features ** -0.5
        self.norm = nn.LayerNorm(features)
        self.norm_ctx = nn.LayerNorm(ctx)
        self.to_q = nn.Linear(features, mid, bias=False)
        self.to_kv = nn.Linear(ctx, 2 * mid, bias=False)
        self.to_out = nn.Linear(mid, features, bias=False)

    def forward(self, x: Tensor, context: Optional[Tensor] = None) -> Tensor:    # x: [B, N, C]
        skip = x
        context = x if context is None else context
        x, context = self.norm(x), self.norm_ctx(context)
        q = self.to_q(x)
        k, v = self.to_kv(context).chunk(2, dim=-1)
        b, n, _ = q.shape
        h = self.num_heads
        q, k, v = (t.reshape(b, t.shape[1], h, -1).transpose(1, 2) for t in (q, k, v))
        sim = torch.einsum("bhnd,bhmd->bhnm", q, k) * self.scale
        attn = sim.softmax(dim=-1)
        out = torch.einsum("bhnm,bhmd->bhnd", attn, v).transpose(1, 2).reshape(b, n, -1)
        return skip + self.to_out(out)


class AttentionItem(nn.Module):
    """A.7 ``AttentionItem``: Packed(Attention) self-attention (d4-d7 in ``exp/model/diffusion.yaml:20``)."""

    def __init__(self, channels: int, head_features: int, num_heads: int):
        super().__init__()
        self.attn = Attention(channels, head_features, num_heads)

    def forward(self, x: Tensor) -> Tensor:
        return self.attn(x.transpose(1, 2)).transpose(1, 2)


class CrossAttentionItem(nn.Module):
    """A.7 ``CrossAttentionItem``: Packed(Attention) with context = CLAP embedding ``[B, M, 512]``
    (``main/module_diffusion.py:67,71`` make M = 1)."""

    def __init__(self, channels: int, head_features: int, num_heads: int, embedding_features: int):
        super().__init__()
        self.attn = Attention(channels, head_features, num_heads, context_features=embedding_features)

    def forward(self, x: Tensor, embedding: Tensor) -> Tensor:
        assert embedding is not None, "CrossAttentionItem requires `embedding`"
        return self.attn(x.transpose(1, 2), embedding).transpose(1, 2)


class ItemGroup(nn.Module):
    """One ``[Resnet, Modulation, InjectChannels?, Attention?, CrossAttention?]`` group (A.1 item list)."""

    def __init__(self, cfg: UNetConfig, d: int):
        super().__init__()
        c = cfg.channels[d]
        self.resnet = ResnetItem(c, cfg.resnet_groups)
        self.mod = ModulationItem(c, cfg.modulation_features)
        self.inject = InjectChannelsItem(c, d, cfg.context_channels[d]) if cfg.context_channels[d] > 0 else None
        self.attn = AttentionItem(c, cfg.attention_features, cfg.attention_heads) if cfg.attentions[d] else None
        self.xattn = (CrossAttentionItem(c, cfg.attention_features, cfg.attention_heads, cfg.embedding_features)
                      if cfg.cross_attentions[d] else None)

    def forward(self, x, features, embedding, channels):
        x = self.resnet(x)
        x = self.mod(x, features)
        if self.inject is not None:
            x = self.inject(x, channels)
        if self.attn is not None:
            x = self.attn(x)
        if self.xattn is not None:
            x = self.xattn(x, embedding)
        return x


class UpsampleNearest(nn.Module):
    """A.6 (N): nn.Upsample(scale_factor=f, mode="nearest") -> Conv1d(k=3, p=1)."""

    def __init__(self, cin: int, cout: int, factor: int):
        super().__init__()
        self.factor = factor
        self.conv = nn.Conv1d(cin, cout, 3, padding=1)

    def forward(self, x: Tensor) -> Tensor:
        if self.factor > 1:
            x = F.interpolate(x, scale_factor=self.factor, mode="nearest")
        return self.conv(x)


class Block(nn.Module):
    """A.5 ``Block``: skip(x, Up(items_up(inner(items_down(Down(x))))), features) with SkipModulate
    ``x + Linear(SiLU(features))[:, :, None] * y``."""

    def __init__(self, cfg: UNetConfig, d: int):
        super().__init__()
        cin = cfg.in_channels if d == 0 else cfg.channels[d - 1]
        c, f = cfg.channels[d], cfg.factors[d]
        self.down = nn.Conv1d(cin, c, kernel_size=f, stride=f)             # A.6 Downsample (patchify; d0: 1x1)
        self.items_down = nn.ModuleList([ItemGroup(cfg, d) for _ in range(cfg.items[d])])
        self.inner = Block(cfg, d + 1) if d + 1 < cfg.depth else None
        self.items_up = nn.ModuleList([ItemGroup(cfg, d) for _ in range(cfg.items[d])])
        if cfg.upsample_mode == "transpose":
            self.up = nn.ConvTranspose1d(c, cin, kernel_size=f, stride=f)  # A.6 (T)
        elif cfg.upsample_mode == "nearest":
            self.up = UpsampleNearest(c, cin, f)                           # A.6 (N)
        else:
            raise ValueError(f"unknown upsample_mode {cfg.upsample_mode!r}")
        self.skip = nn.Linear(cfg.modulation_features, cin)                # SkipModulate / MergeModulate

    def forward(self, x, features, embedding, channels):
        y = self.down(x)
        for it in self.items_down:
            y = it(y, features, embedding, channels)
        if self.inner is not None:
            y = self.inner(y, features, embedding, channels)
        for it in self.items_up:
            y = it(y, features, embedding, channels)
        y = self.up(y)
        s = self.skip(F.silu(features))[:, :, None]
        return x + s * y


# --------------------------------------------------------------------------- plugins (A.3, A.4)


class TimeConditioning(nn.Module):
    """A.3 ``TimeConditioningPlugin``: NumberEmbedder(1024) -> GELU -> Repeat(Linear+GELU, 2) with TIED weights."""

    def __init__(self, features: int, dim: int = 256):
        super().__init__()
        assert dim % 2 == 0
        self.weights = nn.Parameter(torch.randn(dim // 2))                 # LearnedPositionalEmbedding
        self.linear = nn.Linear(dim + 1, features)
        self.mlp = nn.Linear(features, features)                          # same instance applied twice

    def forward(self, time: Tensor) -> Tensor:                            # time: [B]
        t = time[:, None]
        freqs = t * self.weights[None, :] * (2 * math.pi)
        emb = torch.cat([t, freqs.sin(), freqs.cos()], dim=-1)
        h = F.gelu(self.linear(emb))
        h = F.gelu(self.mlp(h))
        h = F.gelu(self.mlp(h))
        return h


class UNetV0(nn.Module):
    """``audio_diffusion_pytorch.UNetV0`` as configured by ``exp/model/diffusion.yaml:13-33``:
    TimeConditioningPlugin(ClassifierFreeGuidancePlugin(XUNet)).  ``forward(x, time, *, embedding,
    embedding_scale, channels)`` is the inner boundary the sampler calls."""

    def __init__(self, cfg: UNetConfig = UNetConfig()):
        super().__init__()
        self.cfg = cfg
        self.time = TimeConditioning(cfg.modulation_features)
        self.fixed_embedding = nn.Embedding(cfg.embedding_max_length, cfg.embedding_features)   # A.4 FixedEmbedding
        self.blocks = Block(cfg, 0)

    def xunet(self, x, features, embedding, channels):
        return self.blocks(x, features, embedding, channels)

    def forward(self, x: Tensor, time: Tensor, *, embedding: Optional[Tensor] = None,
                embedding_scale: float = 1.0, channels: Optional[Sequence[Tensor]] = None) -> Tensor:
        features = self.time(time)
        assert embedding is not None, "ClassifierFreeGuidancePlugin requires `embedding`"       # A.4 assert
        assert embedding.shape[1] <= self.cfg.embedding_max_length, "embedding longer than embedding_max_length"
        if embedding_scale != 1.0:
            # A.4: two SEQUENTIAL passes, then out_m + (out - out_m) * scale
            b, m = embedding.shape[0], embedding.shape[1]
            mask = self.fixed_embedding(torch.arange(m, device=x.device))[None].expand(b, -1, -1)
            out = self.xunet(x, features, embedding, channels)
            out_m = self.xunet(x, features, mask, channels)
            return out_m + (out - out_m) * embedding_scale
        return self.xunet(x, features, embedding, channels)


def count_parameters(m: nn.Module) -> int:
    return sum(p.numel() for p in m.parameters())


@torch.no_grad()
def stress_init_(net: UNetV0, seed: int = 7) -> UNetV0:
    """SURVEY §8(c)-6 "stress init": make norms, attention and the skip/modulation paths non-trivial so parity
    tests do not pass by accident under PyTorch's near-identity default init."""
    g = torch.Generator().manual_seed(seed)
    for name, p in net.named_parameters():
        if name.endswith(("gn1.weight", "gn2.weight", "norm.weight", "norm_ctx.weight")):
            p.copy_(torch.rand(p.shape, generator=g) + 0.5)
        elif name.endswith(("gn1.bias", "gn2.bias", "norm.bias", "norm_ctx.bias")):
            p.copy_((torch.rand(p.shape, generator=g) - 0.5) * 0.5)
        elif name.endswith(("to_q.weight", "to_kv.weight")):
            p.mul_(4.0)
        elif name.endswith("skip.bias"):
            p.add_(1.0)        # skip scale ~ 1 so that deep-level errors are not attenuated on the way up
        elif ".skip." in name or ".mod.linear" in name:
            p.mul_(2.0)
    return net
