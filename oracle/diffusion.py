"""Restatement of ``audio_diffusion_pytorch`` {DiffusionModel, VSampler, VDiffusion, LinearSchedule}
(oracle; test infrastructure only; parity unpinned - see ``oracle/__init__.py``).

Serves the reference call sites ``main/generation.py:77-83`` and ``main/module_diffusion.py:200-206``
(``model.model.sample(x_noisy=, num_steps=, channels=, embedding=, embedding_scale=)``) and the training
step ``main/module_diffusion.py:73-77``.  Semantics: SURVEY.md Appendix A.1-A.2.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from .a_unet import UNetConfig, UNetV0


class LinearSchedule(nn.Module):
    """A.2: sigmas = linspace(start=1, end=0, num_steps + 1)  (``exp/model/diffusion.yaml:27-29`` uses the default)."""

    def __init__(self, start: float = 1.0, end: float = 0.0):
        super().__init__()
        self.start, self.end = start, end

    def forward(self, num_steps: int, device) -> Tensor:
        return torch.linspace(self.start, self.end, num_steps + 1, device=device)


class VSampler(nn.Module):
    """A.2 ``VSampler.forward``: deterministic DDIM-style v-objective sampler, three-line update kept verbatim."""

    def __init__(self, net: nn.Module, schedule: Optional[nn.Module] = None):
        super().__init__()
        self.net = net
        self.schedule = schedule if schedule is not None else LinearSchedule()

    @staticmethod
    def get_alpha_beta(sigmas: Tensor) -> Tuple[Tensor, Tensor]:
        angle = sigmas * math.pi / 2
        return torch.cos(angle), torch.sin(angle)

    @torch.no_grad()
    def forward(self, x_noisy: Tensor, num_steps: int, show_progress: bool = False,
                return_trajectory: bool = False, teacher: Optional[List[Tensor]] = None, **kwargs):
        b = x_noisy.shape[0]
        sigmas = self.schedule(num_steps, device=x_noisy.device).to(x_noisy.dtype)
        sigmas = sigmas[:, None].expand(-1, b)                                  # "i -> i b"
        sigmas_batch = sigmas.reshape(num_steps + 1, b, *([1] * (x_noisy.ndim - 1)))
        alphas, betas = self.get_alpha_beta(sigmas_batch)
        xs, vs = [x_noisy], []
        for i in range(num_steps):
            v_pred = self.net(x_noisy, sigmas[i].contiguous(), **kwargs)
            x_pred = alphas[i] * x_noisy - betas[i] * v_pred
            noise_pred = betas[i] * x_noisy + alphas[i] * v_pred
            x_noisy = alphas[i + 1] * x_pred + betas[i + 1] * noise_pred
            if return_trajectory:
                xs.append(x_noisy)
                vs.append(v_pred)
        if return_trajectory:
            return x_noisy, xs, vs
        return x_noisy


class VDiffusion(nn.Module):
    """A.2 ``VDiffusion.forward`` (training objective; out of the inference hot path, restated for completeness;
    called at ``main/module_diffusion.py:77``)."""

    def __init__(self, net: nn.Module, loss_fn=F.mse_loss):
        super().__init__()
        self.net, self.loss_fn = net, loss_fn

    def forward(self, x: Tensor, generator: Optional[torch.Generator] = None, **kwargs) -> Tensor:
        b = x.shape[0]
        sigmas = torch.rand(b, device=x.device, dtype=x.dtype, generator=generator)
        sb = sigmas.reshape(b, *([1] * (x.ndim - 1)))
        noise = torch.randn(x.shape, device=x.device, dtype=x.dtype, generator=generator)
        alphas, betas = torch.cos(sb * math.pi / 2), torch.sin(sb * math.pi / 2)
        x_noisy = alphas * x + betas * noise
        v_target = alphas * noise - betas * x
        v_pred = self.net(x_noisy, sigmas, **kwargs)
        return self.loss_fn(v_pred, v_target)


class DiffusionModel(nn.Module):
    """A.1 ``DiffusionModel``: ``.net`` (UNetV0), ``.diffusion`` (VDiffusion), ``.sampler`` (VSampler);
    ``forward`` = training loss, ``sample`` = no-grad sampler pass-through - the DROP-IN BOUNDARY
    (``main/generation.py:77-83``)."""

    def __init__(self, cfg: UNetConfig = UNetConfig()):
        super().__init__()
        self.net = UNetV0(cfg)
        self.diffusion = VDiffusion(self.net)
        self.sampler = VSampler(self.net)

    def forward(self, *args, **kwargs) -> Tensor:
        return self.diffusion(*args, **kwargs)

    @torch.no_grad()
    def sample(self, *args, **kwargs) -> Tensor:
        return self.sampler(*args, **kwargs)
