"""Host-side mirror of the reference's sampling interface over the C ABI.

Reference surface kept (same names, argument meaning and error behaviour):

* ``DiffusionModel.sample(x_noisy=, num_steps=, channels=, embedding=, embedding_scale=)``
  - /root/reference/main/generation.py:77-83, main/module_diffusion.py:200-206
* ``DiffusionModel.net(x, time, embedding=, embedding_scale=, channels=)`` - the inner boundary the upstream
  ``VSampler`` calls; ``DiffusionModel.sampler`` - ``VSampler(net)``.
* ``load_state_dict`` - /root/reference/main/generation.py:40-43.

Everything numerical happens in ``libsyncfusion_b200.so``; this file only validates arguments (raising the same
``AssertionError``s as the upstream items), owns the device workspace and forwards pointers + the current CUDA stream.
There is no CPU / PyTorch fallback: importing works anywhere, constructing a model needs a B200 and the built library.
"""
from __future__ import annotations

import ctypes as C
import re
from typing import Dict, Mapping, Optional, Sequence

import torch
from torch import Tensor

from . import _lib
from .config import UNetConfig

_PRECISION = {"fp32": 0, "bf16": 1}
_UPSAMPLE = {"nearest": 0, "transpose": 1}


def flat_param_name(name: str) -> str:
    """Map a (nested) ``state_dict`` key of the U-Net onto the C ABI's flat parameter name.

    ``blocks.inner.inner.items_down.0.resnet.conv1.weight`` -> ``d2.items_down.0.resnet.conv1.weight``;
    ``time.*`` / ``fixed_embedding.weight`` are unchanged; a leading ``net.`` / ``model.net.`` (Lightning
    checkpoint prefix, main/generation.py:43) is dropped.
    """
    name = re.sub(r"^(model\.)?(net\.)", "", name)
    if name.startswith("blocks."):
        rest = name[len("blocks."):]
        d = 0
        while rest.startswith("inner."):
            rest = rest[len("inner."):]
            d += 1
        rest = rest.replace("skip.linear.", "skip.") if rest.startswith("skip.linear.") else rest
        return f"d{d}.{rest}"
    return name


def _c_config(cfg: UNetConfig) -> _lib.SfbUnetConfig:
    c = _lib.SfbUnetConfig()
    c.depth = cfg.depth
    c.in_channels = cfg.in_channels
    for field in ("channels", "factors", "items", "attentions", "cross_attentions", "context_channels"):
        vals = list(getattr(cfg, field))
        assert len(vals) == cfg.depth, f"{field} must have {cfg.depth} entries"
        arr = getattr(c, field)
        for i, v in enumerate(vals):
            arr[i] = int(v)
    c.attention_heads = cfg.attention_heads
    c.attention_features = cfg.attention_features
    c.embedding_features = cfg.embedding_features
    c.embedding_max_length = cfg.embedding_max_length
    c.resnet_groups = cfg.resnet_groups
    c.modulation_features = cfg.modulation_features
    c.upsample_mode = _UPSAMPLE[cfg.upsample_mode]
    c.precision = _PRECISION[cfg.precision]
    return c


class UNetV0:
    """``audio_diffusion_pytorch.UNetV0`` stand-in: ``net(x, time, *, embedding, embedding_scale, channels) -> v``."""

    def __init__(self, cfg: UNetConfig = UNetConfig(), device: "torch.device | str | int" = "cuda"):
        self.cfg = cfg
        self._lib = _lib.load()                      # raises if the CUDA library is missing
        if not torch.cuda.is_available():
            raise _lib.SfbError("syncfusion_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.SfbError(f"device must be CUDA, got {dev}")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self._h = C.c_void_p()
        cc = _c_config(cfg)
        rc = self._lib.sfb_create(C.byref(cc), self.device.index, C.byref(self._h))
        if rc != 0:
            raise _lib.SfbError(f"sfb_create failed with status {rc} (needs an sm_100a GPU)")
        self._finalized = False
        self._ws: Dict[tuple, Tensor] = {}

    # ------------------------------------------------------------------ plumbing
    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.sfb_destroy(h)
            except Exception:
                pass

    def _check(self, rc: int):
        if rc != 0:
            msg = self._lib.sfb_last_error(self._h).decode()
            if rc == -1:
                raise AssertionError(msg)            # upstream raises AssertionError for shape / argument errors
            raise _lib.SfbError(f"[{rc}] {msg}")

    def load_state_dict(self, state_dict: Mapping[str, Tensor], strict: bool = True):
        """Accepts the oracle's / a Lightning checkpoint's nested keys or the flat C-ABI names."""
        assert not self._finalized, "parameters are already finalized"
        for k, v in state_dict.items():
            if not isinstance(v, Tensor):
                continue
            name = flat_param_name(k)
            t = v.detach().to(torch.float32).contiguous().cpu()
            shape = (C.c_int64 * max(t.ndim, 1))(*t.shape)
            self._check(self._lib.sfb_set_param(self._h, name.encode(), C.c_void_p(t.data_ptr()), 0, shape, t.ndim))
        self._check(self._lib.sfb_finalize(self._h))
        self._finalized = True
        return self

    def _workspace(self, B: int, L: int, cfg_on: int, rows: int, M: int = 1) -> Tensor:
        n = C.c_size_t()
        if M == 1 or not hasattr(self._lib, "sfb_workspace_bytes_m"):
            self._check(self._lib.sfb_workspace_bytes(self._h, B, L, cfg_on, rows, C.byref(n)))
        else:
            self._check(self._lib.sfb_workspace_bytes_m(self._h, B, L, cfg_on, rows, M, C.byref(n)))
        key = (B, L, cfg_on, M)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < n.value:
            self._ws.clear()                          # one live workspace at a time
            ws = torch.empty(n.value, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    def _validate(self, x: Tensor, embedding: Optional[Tensor], channels: Optional[Sequence[Tensor]]):
        cfg = self.cfg
        assert self._finalized, "load_state_dict() first"
        assert x.ndim == 3 and x.shape[1] == cfg.in_channels, f"x must be [B, {cfg.in_channels}, L]"
        assert embedding is not None, "ClassifierFreeGuidancePlugin requires `embedding`"
        B, _, L = x.shape
        assert L % cfg.total_factor == 0, f"length {L} must be a multiple of {cfg.total_factor}"
        assert embedding.ndim == 3 and embedding.shape[0] == B and embedding.shape[2] == cfg.embedding_features, \
            f"embedding must be [B, M, {cfg.embedding_features}]"
        assert embedding.shape[1] <= cfg.embedding_max_length, "embedding longer than embedding_max_length"
        assert channels is not None and len(channels) >= cfg.depth, "context `channels` missing for this depth"
        chans = []
        for d in range(cfg.depth):
            want = (B, cfg.context_channels[d], cfg.length_at(L, d))
            assert tuple(channels[d].shape) == want, f"channels[{d}] must be {want}, got {tuple(channels[d].shape)}"
            chans.append(channels[d].to(device=self.device, dtype=torch.float32).contiguous())
        return B, L, chans

    @staticmethod
    def _ptr(t: Optional[Tensor]) -> C.c_void_p:
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()

    # ------------------------------------------------------------------ the two boundaries
    @torch.no_grad()
    def __call__(self, x: Tensor, time: Tensor, *, embedding: Optional[Tensor] = None, embedding_scale: float = 1.0,
                 channels: Optional[Sequence[Tensor]] = None) -> Tensor:
        B, L, chans = self._validate(x, embedding, channels)
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        time = time.to(device=self.device, dtype=torch.float32).reshape(-1).contiguous()
        assert time.numel() == B, "time must be [B]"
        emb = embedding.to(device=self.device, dtype=torch.float32).contiguous()
        cfg_on = int(float(embedding_scale) != 1.0)
        ws = self._workspace(B, L, cfg_on, B, emb.shape[1])
        out = torch.empty_like(x)
        cp = (C.c_void_p * len(chans))(*[c.data_ptr() for c in chans])
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            self._check(self._lib.sfb_unet_forward(self._h, self._ptr(x), self._ptr(time), cp, len(chans), self._ptr(emb),
                                                   emb.shape[1], float(embedding_scale), self._ptr(out), B, L,
                                                   self._ptr(ws), ws.numel(), C.c_void_p(st)))
        return out

    @torch.no_grad()
    def sample(self, x_noisy: Tensor, num_steps: int, *, embedding: Optional[Tensor] = None,
               embedding_scale: float = 1.0, channels: Optional[Sequence[Tensor]] = None,
               return_trajectory: bool = False, teacher: Optional[Tensor] = None):
        """Whole VSampler loop in one C call (the fused fast path ``VSampler.forward`` uses)."""
        B, L, chans = self._validate(x_noisy, embedding, channels)
        x = x_noisy.to(device=self.device, dtype=torch.float32).contiguous()
        emb = embedding.to(device=self.device, dtype=torch.float32).contiguous()
        cfg_on = int(float(embedding_scale) != 1.0)
        ws = self._workspace(B, L, cfg_on, num_steps + 1, emb.shape[1])
        out = torch.empty_like(x)
        tx = tv = None
        if return_trajectory:
            tx = torch.empty((num_steps,) + tuple(x.shape), device=self.device, dtype=torch.float32)
            tv = torch.empty_like(tx)
        if teacher is not None:
            teacher = teacher.to(device=self.device, dtype=torch.float32).contiguous()
            assert tuple(teacher.shape) == (num_steps,) + tuple(x.shape)
        cp = (C.c_void_p * len(chans))(*[c.data_ptr() for c in chans])
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            self._check(self._lib.sfb_sample(self._h, self._ptr(x), int(num_steps), cp, len(chans), self._ptr(emb),
                                             emb.shape[1], float(embedding_scale), self._ptr(out), self._ptr(tx),
                                             self._ptr(tv), self._ptr(teacher), B, L, self._ptr(ws), ws.numel(),
                                             C.c_void_p(st)))
        if return_trajectory:
            return out, tx, tv
        return out

    @property
    def last_launch_count(self) -> int:
        return int(self._lib.sfb_last_launch_count(self._h))

    # ------------------------------------------------------------------ test hooks
    def debug_ops(self, B: int, L: int, cfg_on: int, M: int = 1):
        ws = self._workspace(B, L, cfg_on, B, M)
        n = (self._lib.sfb_dbg_plan_size(self._h, B, L, cfg_on, self._ptr(ws), ws.numel()) if M == 1 else
             self._lib.sfb_dbg_plan_size_m(self._h, B, L, cfg_on, M, self._ptr(ws), ws.numel()))
        if n < 0:
            self._check(n)
        ops = []
        buf = C.create_string_buffer(256)
        for i in range(n):
            self._check(self._lib.sfb_dbg_op_info(self._h, i, buf, 256))
            kind, depth, stack, item, off, nbytes, rows, cols, dt, ck = buf.value.decode().split()
            ops.append(dict(kind=kind, depth=int(depth), stack=int(stack), item=int(item), off=int(off),
                            nbytes=int(nbytes), rows=int(rows), cols=int(cols), dtype=int(dt), ck=ck))
        return ops, ws

    def profile(self, enable: bool):
        """Bracket every launch of subsequent U-Net evaluations with CUDA events (see ``profile_report``)."""
        self._check(self._lib.sfb_dbg_profile(self._h, int(enable)))

    def profile_report(self):
        """Per-op device time / algorithmic flops / bytes of the most recent profiled evaluation (sync first)."""
        buf = C.create_string_buffer(1 << 17)
        self._check(self._lib.sfb_dbg_profile_report(self._h, buf, len(buf)))
        rows = []
        for line in buf.value.decode().splitlines():
            i, kind, depth, stack, item, ms, flops, nbytes = line.split()
            rows.append(dict(index=int(i), kind=kind, depth=int(depth), stack=int(stack), item=int(item), ms=float(ms),
                             flops=float(flops), bytes=float(nbytes)))
        return rows

    def wait_log(self) -> str:
        """Decoded device barrier-wait timeout log ('' if none); readable even after the CUDA context is lost."""
        if not hasattr(self._lib, "sfb_dbg_wait_log"):
            return ""
        buf = C.create_string_buffer(1 << 16)
        self._lib.sfb_dbg_wait_log(self._h, buf, len(buf))
        return buf.value.decode(errors="replace")

    def debug_set_grid_limit(self, max_ctas: int):
        """Persistent kernels use at most ``max_ctas`` CTAs (0: one per SM) - more tiles per CTA for the ring tests."""
        self._check(self._lib.sfb_dbg_set_grid_limit(self._h, int(max_ctas)))

    def debug_set_op_limit(self, n: int):
        self._check(self._lib.sfb_dbg_set_op_limit(self._h, int(n)))


class VSampler:
    """``audio_diffusion_pytorch.VSampler`` stand-in: ``forward(x_noisy, num_steps, show_progress=False, **kwargs)``."""

    def __init__(self, net: UNetV0):
        self.net = net

    @torch.no_grad()
    def forward(self, x_noisy: Tensor, num_steps: int, show_progress: bool = False, **kwargs) -> Tensor:
        return self.net.sample(x_noisy, num_steps, **kwargs)

    __call__ = forward


class DiffusionModel(torch.nn.Module):
    """``audio_diffusion_pytorch.DiffusionModel`` stand-in for the sampling path (exp/model/diffusion.yaml:11-33).

    ``model.sample(x_noisy=noise, num_steps=N, channels=y_latent['xs'][2:-1], embedding=z_latent,
    embedding_scale=s)`` works exactly as at main/generation.py:77-83.  It is an ``nn.Module`` WITHOUT parameters (the
    weights live re-packed inside ``libsyncfusion_b200.so``), so it can sit where the reference keeps it - as
    ``Model.model`` of the Lightning module (main/module_diffusion.py:39) - and the reference's own calls reach it:

    * ``Model.load_state_dict(checkpoint['state_dict'])`` (main/generation.py:42-43) recurses into
      ``_load_from_state_dict`` below, which consumes every ``model.*`` key (nested oracle / checkpoint names or flat
      C-ABI names) and reports missing / unexpected ones like any module;
    * ``Model.to(device)`` (main/generation.py:44) moves the zero-size anchor buffer; the C engine is created on that
      device at the first use (or eagerly when a CUDA ``device`` is given to the constructor);
    * ``list(Model.model.parameters())`` (main/module_diffusion.py:55) is empty.

    Training (``forward`` = VDiffusion loss, main/module_diffusion.py:77) is outside the hot path and raises.
    """

    def __init__(self, cfg: UNetConfig = UNetConfig(), device: "torch.device | str | int | None" = None):
        super().__init__()
        self.cfg = cfg
        self.register_buffer("_anchor", torch.empty(0), persistent=False)     # follows .to(device) / .cuda()
        self._staged: Dict[str, Tensor] = {}        # flat name -> fp32 CPU tensor, until the engine is finalized
        self._net: Optional[UNetV0] = None
        self._sampler: Optional[VSampler] = None
        if device is not None:
            dev = torch.device(device)
            if dev.type != "cuda":
                raise _lib.SfbError(f"device must be CUDA, got {dev}")
            if not torch.cuda.is_available():
                raise _lib.SfbError("syncfusion_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            self._anchor = self._anchor.to(torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device()))
            self._net = UNetV0(cfg, self._anchor.device)      # eager: constructing with a device needs the B200 + library
            self._sampler = VSampler(self._net)

    # ------------------------------------------------------------------ engine life cycle
    @property
    def device(self) -> torch.device:
        return self._anchor.device

    def _engine(self) -> UNetV0:
        """The C engine on this module's device: created at first use, fed with the staged parameters and finalized."""
        dev = self._anchor.device
        if dev.type != "cuda":
            raise _lib.SfbError("syncfusion_b200.DiffusionModel is on the CPU: move it with .to('cuda') first "
                                "(the sampling path is CUDA-only, there is no CPU fallback)")
        if self._net is not None and self._net.device != dev:
            if not self._staged:
                raise _lib.SfbError(f"the engine was finalized on {self._net.device}; re-load the state dict to move it to {dev}")
            self._net = None
        if self._net is None:
            self._net = UNetV0(self.cfg, dev)
            self._sampler = VSampler(self._net)
        if not self._net._finalized:
            if not self._staged:
                raise AssertionError("load_state_dict() first")
            self._net.load_state_dict(self._staged)
            self._staged = {}                          # the library holds the re-packed copies now
        return self._net

    @property
    def net(self) -> UNetV0:
        return self._engine()

    @property
    def sampler(self) -> VSampler:
        self._engine()
        return self._sampler

    # ------------------------------------------------------------------ state dict (main/generation.py:40-43)
    def _stage(self, state_dict: Mapping[str, Tensor], prefix: str, strict: bool, missing: list, unexpected: list, errors: list):
        from .synth import param_shapes
        want = param_shapes(self.cfg)
        got: Dict[str, Tensor] = {}
        mine = [(k, v) for k, v in state_dict.items() if k.startswith(prefix) and isinstance(v, Tensor)]
        unknown, bad = [], []
        for k, v in mine:
            name = flat_param_name(k[len(prefix):])
            if name not in want:
                unknown.append(k)
                continue
            if tuple(v.shape) != want[name]:
                bad.append(f"size mismatch for {k}: checkpoint {tuple(v.shape)} vs model {want[name]}")
                continue
            got[name] = v
        if len(unknown) * 2 >= len(mine) and len(mine) > 0:
            # names this package does not know (a checkpoint written by the upstream classes): map the
            # tensors by registration order and shape (checkpoint.py); on failure the name-based report below stands
            from .checkpoint import structural_key_map
            try:
                m = structural_key_map([(k, tuple(v.shape)) for k, v in mine], self.cfg, dict(mine))
                got = {m[k]: v for k, v in mine if not m[k].endswith("#alias")}
                unknown, bad = [], []
                self.key_map = {k: m[k] for k, _ in mine}
            except KeyError as e:
                self.key_map_error = str(e)
                if bad == [] and strict:
                    errors.append(f"structural key mapping failed: {e}")
        unexpected.extend(unknown)
        errors.extend(bad)
        got = {n: v.detach().to(torch.float32).cpu().contiguous() for n, v in got.items()}
        n_missing = 0
        for name in want:
            if name not in got:
                missing.append(prefix + name)
                n_missing += 1
        if not errors and (n_missing == 0 or not strict):
            if self._net is not None and self._net._finalized:
                self._net = None                       # re-load: a fresh engine is built from the new weights
            self._staged = got

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        # called by nn.Module.load_state_dict of THIS module or of any parent (the reference's Lightning Model)
        self._stage(state_dict, prefix, strict, missing_keys, unexpected_keys, error_msgs)

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        # the weights are owned by the C library in re-packed form; nothing to export (and nothing for a parent to save)
        return destination if destination is not None else {}

    # ------------------------------------------------------------------ the reference surface
    @torch.no_grad()
    def sample(self, *args, **kwargs) -> Tensor:
        return self.sampler(*args, **kwargs)

    def forward(self, *args, **kwargs):
        raise NotImplementedError("the VDiffusion training objective is out of scope of the B200 sampling path")


def hydra_config(net_t=None, diffusion_t=None, sampler_t=None, use_embedding_cfg: bool = True, precision: str = "bf16",
                 upsample_mode: str = "nearest", **kw) -> UNetConfig:
    """``UNetConfig`` from the keyword arguments Hydra passes for ``exp/model/diffusion.yaml:11-33`` (the partials
    ``net_t`` / ``diffusion_t`` / ``sampler_t`` select upstream classes this package replaces and are ignored)."""
    del net_t, diffusion_t, sampler_t
    fields = {f for f in UNetConfig.__dataclass_fields__}
    unknown = sorted(set(kw) - fields)
    if unknown:
        raise TypeError(f"unsupported DiffusionModel arguments for the B200 sampling path: {unknown}")
    seq = ("channels", "factors", "items", "attentions", "cross_attentions", "context_channels")
    kw = {k: (tuple(int(x) for x in v) if k in seq else v) for k, v in kw.items()}
    return UNetConfig(use_embedding_cfg=bool(use_embedding_cfg), precision=precision, upsample_mode=upsample_mode, **kw)


def hydra_target(device: "torch.device | str | int | None" = None, **kw) -> DiffusionModel:
    """Drop-in ``_target_`` for ``model.model`` of exp/model/diffusion.yaml."""
    return DiffusionModel(hydra_config(**kw), device)

