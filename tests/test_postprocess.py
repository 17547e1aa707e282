"""f-2 oracle pin (CPU): oracle/postprocess.py against the real dependency, torchaudio.functional.resample, and the
reference's mask / crop order (main/generation.py:85-98); committed golden vector."""
import os

import numpy as np
import pytest
import torch

from oracle.postprocess import postprocess, resample, sinc_resample_table

torchaudio = pytest.importorskip("torchaudio")
GOLD = os.path.join(os.path.dirname(__file__), "golden", "postprocess_golden.npz")


@pytest.mark.parametrize("orig,new,n", [(48000, 22050, 96000), (48000, 22050, 4097), (48000, 16000, 5000), (44100, 48000, 3001), (48000, 44100, 2048)])
def test_oracle_resample_matches_torchaudio(orig, new, n):
    g = torch.Generator().manual_seed(n)
    x = torch.randn(2, 1, n, generator=g)
    ref = torchaudio.functional.resample(x, orig_freq=orig, new_freq=new).numpy()
    got = resample(x.numpy(), orig, new)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())


def test_table_matches_torchaudio_kernel():
    import math
    import torchaudio.functional.functional as F
    k, width = F._get_sinc_resample_kernel(48000, 22050, math.gcd(48000, 22050), dtype=torch.float32)
    t, w, orig, new = sinc_resample_table(48000, 22050)
    assert (w, orig, new) == (width, 320, 147) and t.shape == (147, 2 * width + 320)
    assert np.abs(k[:, 0].numpy() - t).max() < 3e-7


def _reference_lines(gen, y, cut_prefix, cut_length, sample_rate, downsample_rate):
    """main/generation.py:85-98 as the reference executes it (torch + torchaudio), one clip at a time."""
    gen = gen.clone()
    outs = []
    for i in range(gen.shape[0]):
        if cut_prefix:
            first_onset = torch.nonzero(y[i][0]).squeeze(-1)[0]
            gen[i, :, :first_onset] = 0.
        if downsample_rate:
            outs.append(torchaudio.functional.resample(gen[i, :, :cut_length].cpu(), orig_freq=sample_rate, new_freq=downsample_rate))
        else:
            outs.append(gen[i, :, :cut_length].cpu())
    return torch.stack(outs)


@pytest.mark.parametrize("cut_prefix", [False, True])
@pytest.mark.parametrize("down", [None, 22050])
def test_oracle_postprocess_matches_reference_sequence(cut_prefix, down):
    g = torch.Generator().manual_seed(3)
    B, L, cut = 3, 8192, 6000
    gen = torch.randn(B, 1, L, generator=g)
    y = torch.zeros(B, 1, L)
    for b in range(B):
        y[b, 0, torch.randint(50, L, (4,), generator=g)] = 1.0
    ref = _reference_lines(gen, y, cut_prefix, cut, 48000, down).numpy()
    got = postprocess(gen.numpy(), y.numpy(), cut_prefix=cut_prefix, cut_length=cut, sample_rate=48000, downsample_rate=down)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())
    if cut_prefix:
        with pytest.raises(IndexError):
            postprocess(gen.numpy(), np.zeros_like(y.numpy()), cut_prefix=True, cut_length=cut, downsample_rate=down)


def test_golden_vector():
    z = np.load(GOLD)
    got = postprocess(z["gen"], z["onsets"], cut_prefix=True, cut_length=int(z["cut_length"]), sample_rate=48000, downsample_rate=22050)
    assert got.shape == z["out"].shape
    assert np.abs(got - z["out"]).max() <= 2e-6
