"""Generation post-processing on the GPU: the step right after ``model.model.sample`` in
/root/reference/main/generation.py:85-98, for the whole batch in one kernel, so that only the cropped / resampled
audio crosses PCIe.  Host-side mirror over ``sfb_postprocess``; no CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import Tensor

from . import _lib


def postprocess(gen: Tensor, onsets: Optional[Tensor] = None, cut_prefix: bool = False, cut_length: Optional[int] = None,
                sample_rate: int = 48000, downsample_rate: Optional[int] = None) -> Tensor:
    """``gen`` [B, 1, L] (CUDA) -> [B, 1, T]: what the reference hands to ``torchaudio.save`` for every clip.

    Same keyword meaning as ``generate_dataset`` (main/generation.py:24-28): ``cut_prefix`` zeroes everything before the
    clip's first onset (``onsets`` = the onset tracks ``y`` [B, 1, L]); ``cut_length`` crops (default: L);
    ``downsample_rate`` resamples ``sample_rate -> downsample_rate`` exactly as ``torchaudio.functional.resample``.
    Raises ``IndexError`` like the reference when ``cut_prefix`` is set and a clip has no onset."""
    lib = _lib.load()
    if not gen.is_cuda:
        raise _lib.SfbError("postprocess needs CUDA tensors (there is no CPU fallback; the oracle lives in oracle/postprocess.py)")
    assert gen.ndim == 3 and gen.shape[1] == 1, "gen must be [B, 1, L]"
    B, _, L = gen.shape
    cut = L if not cut_length else int(cut_length)
    assert 0 < cut <= L, "cut_length must be in (0, L]"
    g = gen.to(torch.float32).contiguous()
    y = None
    first = None
    if cut_prefix:
        assert onsets is not None and tuple(onsets.shape) == (B, 1, L), "cut_prefix needs the onset tracks y [B, 1, L]"
        y = onsets.to(device=gen.device, dtype=torch.float32).contiguous()
        first = torch.empty(B, dtype=torch.int32, device=gen.device)
    new = int(downsample_rate) if downsample_rate else 0
    T = int(lib.sfb_postprocess_out_len(cut, int(sample_rate), new))
    out = torch.empty(B, 1, T, device=gen.device, dtype=torch.float32)
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()   # noqa: E731
    with torch.cuda.device(gen.device):
        st = torch.cuda.current_stream().cuda_stream
        rc = lib.sfb_postprocess(gen.device.index, ptr(g), ptr(y), B, L, cut, int(sample_rate), new, ptr(out), T, ptr(first), C.c_void_p(st))
    if rc != 0:
        raise _lib.SfbError(f"sfb_postprocess failed with status {rc}")
    if cut_prefix and bool((first >= L).any().item()):
        raise IndexError("index 0 is out of bounds for dimension 0 with size 0 (a clip has no onset: torch.nonzero(y[i][0])[0])")
    return out
