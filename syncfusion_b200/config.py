"""Frozen mirror of /root/reference/exp/model/diffusion.yaml:15-33 (plus the two knobs the port adds)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence


@dataclass(frozen=True)
class UNetConfig:
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
    resnet_groups: int = 8
    modulation_features: int = 1024
    upsample_mode: str = "nearest"      # "nearest" | "transpose"  
    precision: str = "bf16"             # "bf16" (bf16 operands, fp32 residual stream) | "fp32" (TF32 MMA, fp32 storage)

    @property
    def depth(self) -> int:
        return len(self.channels)

    @property
    def total_factor(self) -> int:
        f = 1
        for x in self.factors:
            f *= x
        return f

    def length_at(self, length: int, d: int) -> int:
        f = 1
        for i in range(d + 1):
            f *= self.factors[i]
        return length // f
