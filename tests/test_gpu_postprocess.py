"""f-2 on the GPU (needs a B200): sfb_postprocess (mask before the first onset + crop + 48 k -> 22.05 k polyphase
resample in one kernel; main/generation.py:85-98) against the oracle, against torchaudio itself where it is installed,
and against the committed golden vector.  Tolerance: the kernel sums each output's 348 taps in a different order than
torchaudio's conv1d, so 2e-6 absolute per unit signal scale (fp32 round-off), not bit equality."""
import os

import numpy as np
import pytest
import torch

from oracle.postprocess import postprocess as oracle_postprocess

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "postprocess_golden.npz")


def _case(B, L, seed):
    g = torch.Generator().manual_seed(seed)
    gen = torch.randn(B, 1, L, generator=g)
    y = torch.zeros(B, 1, L)
    for b in range(B):
        y[b, 0, torch.randint(b, L, (3,), generator=g)] = 1.0
    return gen, y


@pytest.mark.parametrize("cut_prefix", [False, True])
@pytest.mark.parametrize("B,L,cut,sr,down", [
    (16, 262144, 96000, 48000, 22050),      # exp/evaluate_gh_gen.yaml: crop to 96000, 48 k -> 22.05 k
    (3, 8192, 6000, 48000, 22050), (2, 4097 * 2, 4097, 48000, 16000), (2, 4096, 3001, 44100, 48000), (1, 2048, 2048, 48000, 44100),
    (2, 16384, 9000, 48000, None), (2, 1000, None, 48000, None),
])
def test_postprocess_matches_oracle(cuda_device, B, L, cut, sr, down, cut_prefix):
    import syncfusion_b200 as sf
    gen, y = _case(B, L, L + (cut or 0))
    ref = oracle_postprocess(gen.numpy(), y.numpy(), cut_prefix=cut_prefix, cut_length=cut, sample_rate=sr, downsample_rate=down)
    out = sf.postprocess(gen.to(cuda_device), y.to(cuda_device), cut_prefix=cut_prefix, cut_length=cut, sample_rate=sr, downsample_rate=down)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())
    if down is None:
        assert np.array_equal(got, ref)          # mask + crop only: bit exact


def test_postprocess_matches_torchaudio_and_golden(cuda_device):
    import syncfusion_b200 as sf
    z = np.load(GOLD)
    out = sf.postprocess(torch.from_numpy(z["gen"]).to(cuda_device), torch.from_numpy(z["onsets"]).to(cuda_device), cut_prefix=True,
                         cut_length=int(z["cut_length"]), sample_rate=48000, downsample_rate=22050).cpu().numpy()
    assert out.shape == z["out"].shape and np.abs(out - z["out"]).max() <= 2e-6
    torchaudio = pytest.importorskip("torchaudio")
    gen, _ = _case(2, 96000, 5)
    ref = torchaudio.functional.resample(gen, orig_freq=48000, new_freq=22050).numpy()
    got = sf.postprocess(gen.to(cuda_device), sample_rate=48000, downsample_rate=22050).cpu().numpy()
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()


def test_postprocess_errors_mirror_the_reference(cuda_device):
    import syncfusion_b200 as sf
    gen, y = _case(2, 4096, 1)
    y[1] = 0.0                                               # a clip without an onset: torch.nonzero(...)[0] raises IndexError
    with pytest.raises(IndexError):
        sf.postprocess(gen.to(cuda_device), y.to(cuda_device), cut_prefix=True, downsample_rate=22050)
    with pytest.raises(Exception):
        sf.postprocess(gen, y, cut_prefix=True)              # CPU tensors: no CPU fallback
