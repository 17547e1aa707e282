"""Shared helpers for the parity tests: seeded synthetic inputs in the shapes of exp/train_diffusion_gh.yaml."""
from __future__ import annotations

import torch

from oracle import DiffusionModel as OracleModel, Encoder1d, UNetConfig as OracleConfig, stress_init_

SMALL = dict(channels=(8, 32, 64, 128), factors=(1, 4, 4, 2), items=(1, 2, 1, 2), attentions=(0, 0, 0, 1),
             cross_attentions=(1, 1, 1, 1), context_channels=(2, 8, 16, 32))


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def make_oracle(cfg_kwargs=None, seed=0, stress=False, upsample_mode="nearest"):
    torch.manual_seed(seed)
    cfg = OracleConfig(upsample_mode=upsample_mode, **(cfg_kwargs or {}))
    m = OracleModel(cfg).eval()
    if stress:
        stress_init_(m.net)
    return m


def make_encoder(cfg, seed=1):
    torch.manual_seed(seed)
    ctx = list(cfg.context_channels)
    mult = [1] + [c // 2 for c in ctx]
    return Encoder1d(in_channels=1, channels=2, multipliers=mult, factors=list(cfg.factors),
                     num_blocks=[2] * len(ctx), resnet_groups=2, patch_size=1).eval()


@torch.no_grad()
def make_inputs(cfg, B, L, seed=12345, encoder=None):
    """noise ~ N(0,1) [B,1,L] (main/generation.py:69); onset impulse track (main/dataset_diffusion.py:66-72) ->
    Encoder1d pyramid xs[2:-1] (main/generation.py:71,80); unit-norm 512-d 'CLAP' embedding [B,1,512]."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 1, L, generator=g)
    y = torch.zeros(B, 1, L)
    for b in range(B):
        k = int(torch.randint(4, 17, (1,), generator=g))
        pos = torch.randint(0, L, (k,), generator=g)
        y[b, 0, pos] = 1.0
    enc = encoder if encoder is not None else make_encoder(cfg)
    _, info = enc(y, with_info=True)
    channels = [c.contiguous() for c in info["xs"][2:-1]]
    e = torch.randn(B, 1, cfg.embedding_features, generator=g)
    e = e / e.norm(dim=-1, keepdim=True)
    return x, channels, e
