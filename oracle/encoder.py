"""Restatement of ``audio_encoders_pytorch.Encoder1d`` (oracle; test infrastructure only; parity unpinned).

The onset encoder runs ONCE per batch, outside the sampling loop (``main/generation.py:71``,
``main/module_diffusion.py:196``); its ``info['xs'][2:-1]`` pyramid is the hot path's ``channels`` input.
Configured by ``exp/model/diffusion.yaml:35-43``.  Semantics: SURVEY.md Appendix A.8.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor


class ConvBlock1d(nn.Module):
    def __init__(self, cin: int, cout: int, groups: int):
        super().__init__()
        self.groupnorm = nn.GroupNorm(groups, cin)
        self.project = nn.Conv1d(cin, cout, 3, padding=1)

    def forward(self, x: Tensor) -> Tensor:
        return self.project(F.silu(self.groupnorm(x)))


class ResnetBlock1d(nn.Module):
    def __init__(self, cin: int, cout: int, groups: int):
        super().__init__()
        self.block1 = ConvBlock1d(cin, cout, groups if cin % groups == 0 else 1)
        self.block2 = ConvBlock1d(cout, cout, groups)
        self.to_out = nn.Conv1d(cin, cout, 1) if cin != cout else nn.Identity()

    def forward(self, x: Tensor) -> Tensor:
        return self.block2(self.block1(x)) + self.to_out(x)


class DownsampleBlock1d(nn.Module):
    def __init__(self, cin: int, cout: int, factor: int, num_blocks: int, groups: int):
        super().__init__()
        self.downsample = nn.Conv1d(cin, cout, kernel_size=2 * factor + 1, stride=factor, padding=factor)
        self.blocks = nn.ModuleList([ResnetBlock1d(cout, cout, groups) for _ in range(num_blocks)])

    def forward(self, x: Tensor) -> Tensor:
        x = self.downsample(x)
        for b in self.blocks:
            x = b(x)
        return x


class Encoder1d(nn.Module):
    """``forward(x, with_info=True) -> (z, {'xs': [x, to_in(x), ds_0 .. ds_7, to_out(z)]})`` - 11 entries, so
    ``xs[2:-1]`` are the 8 pyramid levels consumed at ``main/generation.py:80``."""

    def __init__(self, in_channels: int = 1, channels: int = 2,
                 multipliers: Sequence[int] = (1, 1, 4, 8, 16, 32, 64, 128, 128),
                 factors: Sequence[int] = (1, 4, 4, 4, 2, 2, 2, 2),
                 num_blocks: Sequence[int] = (2,) * 8, resnet_groups: int = 2, patch_size: int = 1):
        super().__init__()
        assert patch_size == 1, "exp/model/diffusion.yaml:43 uses patch_size 1 (no reshape)"
        assert len(multipliers) == len(factors) + 1 == len(num_blocks) + 1
        self.to_in = ResnetBlock1d(in_channels, channels * multipliers[0], 1)     # Patcher
        self.downsamples = nn.ModuleList([
            DownsampleBlock1d(channels * multipliers[i], channels * multipliers[i + 1], factors[i],
                              num_blocks[i], resnet_groups)
            for i in range(len(factors))])

    def forward(self, x: Tensor, with_info: bool = False):
        xs = [x]
        x = self.to_in(x)
        xs.append(x)
        for ds in self.downsamples:
            x = ds(x)
            xs.append(x)
        xs.append(x)                                                             # to_out = Identity
        return (x, {"xs": xs}) if with_info else x
