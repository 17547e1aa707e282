"""Onset encoder on the GPU: host-side mirror of ``audio_encoders_pytorch.Encoder1d`` over the C ABI.

Same constructor keywords as exp/model/diffusion.yaml:35-43 and the same call as main/generation.py:71 /
main/module_diffusion.py:196::

    _, y_latent = model.onsets_encoder(y, with_info=True)        # y_latent['xs'][2:-1] -> the sampler's `channels`

An ``nn.Module`` without parameters (the weights live in ``libsyncfusion_b200.so``): a parent's ``load_state_dict`` reaches
``_load_from_state_dict`` with the upstream key names (``to_in.block1.groupnorm.weight``, ``downsamples.3.blocks.1...``),
``.to(device)`` moves the anchor buffer and the engine is created on that device at first use.  Inference only; no CPU
fallback (the oracle restatement lives in ``oracle/encoder.py``).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib


def encoder_param_shapes(in_channels: int, channels: int, multipliers: Sequence[int], factors: Sequence[int],
                         num_blocks: Sequence[int]) -> Dict[str, Tuple[int, ...]]:
    sh: Dict[str, Tuple[int, ...]] = {}

    def block(pre: str, cin: int, cout: int):
        sh[pre + "block1.groupnorm.weight"] = sh[pre + "block1.groupnorm.bias"] = (cin,)
        sh[pre + "block1.project.weight"], sh[pre + "block1.project.bias"] = (cout, cin, 3), (cout,)
        sh[pre + "block2.groupnorm.weight"] = sh[pre + "block2.groupnorm.bias"] = (cout,)
        sh[pre + "block2.project.weight"], sh[pre + "block2.project.bias"] = (cout, cout, 3), (cout,)
        if cin != cout:
            sh[pre + "to_out.weight"], sh[pre + "to_out.bias"] = (cout, cin, 1), (cout,)

    block("to_in.", in_channels, channels * multipliers[0])
    for i, f in enumerate(factors):
        cin, cout = channels * multipliers[i], channels * multipliers[i + 1]
        sh[f"downsamples.{i}.downsample.weight"], sh[f"downsamples.{i}.downsample.bias"] = (cout, cin, 2 * f + 1), (cout,)
        for j in range(num_blocks[i]):
            block(f"downsamples.{i}.blocks.{j}.", cout, cout)
    return sh


class Encoder1d(torch.nn.Module):
    """``forward(x, with_info=False) -> z`` or ``(z, {'xs': [x, to_in(x), level_0 .. level_{n-1}, z]})``."""

    def __init__(self, in_channels: int = 1, channels: int = 2, multipliers: Sequence[int] = (1, 1, 4, 8, 16, 32, 64, 128, 128),
                 factors: Sequence[int] = (1, 4, 4, 4, 2, 2, 2, 2), num_blocks: Sequence[int] = (2,) * 8, resnet_groups: int = 2,
                 patch_size: int = 1, device: "torch.device | str | int | None" = None):
        super().__init__()
        assert patch_size == 1, "exp/model/diffusion.yaml:43 uses patch_size 1"
        assert len(multipliers) == len(factors) + 1 == len(num_blocks) + 1
        self.in_channels, self.channels = int(in_channels), int(channels)
        self.multipliers, self.factors, self.num_blocks = tuple(int(m) for m in multipliers), tuple(int(f) for f in factors), tuple(int(n) for n in num_blocks)
        self.resnet_groups, self.patch_size = int(resnet_groups), int(patch_size)
        self.register_buffer("_anchor", torch.empty(0), persistent=False)
        self._staged: Dict[str, Tensor] = {}
        self._h: Optional[C.c_void_p] = None
        self._h_device: Optional[torch.device] = None
        self._ws: Optional[Tensor] = None
        self._lib = None
        if device is not None:
            self._anchor = self._anchor.to(torch.device(device))

    @property
    def device(self) -> torch.device:
        return self._anchor.device

    def _shapes(self):
        return encoder_param_shapes(self.in_channels, self.channels, self.multipliers, self.factors, self.num_blocks)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        want = self._shapes()
        got: Dict[str, Tensor] = {}
        for k, v in state_dict.items():
            if not k.startswith(prefix) or not isinstance(v, Tensor):
                continue
            name = k[len(prefix):]
            if name not in want:
                unexpected_keys.append(k)
            elif tuple(v.shape) != want[name]:
                error_msgs.append(f"size mismatch for {k}: checkpoint {tuple(v.shape)} vs model {want[name]}")
            else:
                got[name] = v.detach().to(torch.float32).cpu().contiguous()
        missing = [prefix + n for n in want if n not in got]
        missing_keys.extend(missing)
        if not error_msgs and (not missing or not strict):
            self._staged = got
            self._destroy()

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        return destination if destination is not None else {}

    def _destroy(self):
        if self._h is not None and self._lib is not None:
            self._lib.sfb_encoder_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:   # noqa: BLE001
            pass

    def _check(self, rc: int):
        if rc != 0:
            msg = self._lib.sfb_encoder_last_error(self._h).decode() if self._h is not None else ""
            if rc == -1:
                raise AssertionError(msg)
            raise _lib.SfbError(f"[{rc}] {msg}")

    def _engine(self):
        dev = self._anchor.device
        if dev.type != "cuda":
            raise _lib.SfbError("syncfusion_b200.Encoder1d is on the CPU: move it with .to('cuda') first (no CPU fallback)")
        if self._h is not None and self._h_device == dev:
            return
        if not self._staged:
            raise AssertionError("load_state_dict() first")
        self._lib = _lib.load()
        self._destroy()
        cfg = _lib.SfbEncoderConfig()
        cfg.in_channels, cfg.channels, cfg.n_levels = self.in_channels, self.channels, len(self.factors)
        for i, m in enumerate(self.multipliers):
            cfg.multipliers[i] = m
        for i, (f, n) in enumerate(zip(self.factors, self.num_blocks)):
            cfg.factors[i], cfg.num_blocks[i] = f, n
        cfg.resnet_groups, cfg.patch_size = self.resnet_groups, self.patch_size
        h = C.c_void_p()
        rc = self._lib.sfb_encoder_create(C.byref(cfg), dev.index if dev.index is not None else torch.cuda.current_device(), C.byref(h))
        if rc != 0:
            raise _lib.SfbError(f"sfb_encoder_create failed with status {rc}")
        self._h, self._h_device = h, dev
        for name, t in self._staged.items():
            shape = (C.c_int64 * max(t.ndim, 1))(*t.shape)
            self._check(self._lib.sfb_encoder_set_param(self._h, name.encode(), C.c_void_p(t.data_ptr()), shape, t.ndim))
        self._check(self._lib.sfb_encoder_finalize(self._h))

    @torch.no_grad()
    def forward(self, x: Tensor, with_info: bool = False):
        self._engine()
        assert x.ndim == 3 and x.shape[1] == self.in_channels, f"x must be [B, {self.in_channels}, L]"
        dev = self._anchor.device
        x = x.to(device=dev, dtype=torch.float32).contiguous()
        B, _, L = x.shape
        n = C.c_size_t()
        self._check(self._lib.sfb_encoder_workspace_bytes(self._h, B, L, C.byref(n)))
        if self._ws is None or self._ws.numel() < n.value or self._ws.device != dev:
            self._ws = torch.empty(n.value, dtype=torch.uint8, device=dev)
        outs: List[Tensor] = [torch.empty(B, self.channels * self.multipliers[0], L, device=dev)]
        for i in range(len(self.factors)):
            Li = int(self._lib.sfb_encoder_level_length(self._h, L, i))
            outs.append(torch.empty(B, self.channels * self.multipliers[i + 1], Li, device=dev))
        ptrs = (C.c_void_p * len(outs))(*[t.data_ptr() for t in outs])
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream().cuda_stream
            self._check(self._lib.sfb_encoder_forward(self._h, C.c_void_p(x.data_ptr()), B, L, ptrs, len(outs),
                                                      C.c_void_p(self._ws.data_ptr()), self._ws.numel(), C.c_void_p(st)))
        z = outs[-1]
        return (z, {"xs": [x] + outs + [z]}) if with_info else z
