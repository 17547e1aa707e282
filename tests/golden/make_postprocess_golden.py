"""Generates tests/golden/postprocess_golden.npz with the REAL dependency of the reference's post-processing
(torchaudio.functional.resample, main/generation.py:85-98):  python tests/golden/make_postprocess_golden.py"""
import os

import numpy as np
import torch
import torchaudio

g = torch.Generator().manual_seed(20261017)
B, L, cut = 2, 12000, 9600
gen = torch.randn(B, 1, L, generator=g)
y = torch.zeros(B, 1, L)
y[0, 0, [700, 3000, 9000]] = 1.0
y[1, 0, [11, 5000]] = 1.0
work = gen.clone()
outs = []
for i in range(B):
    first_onset = torch.nonzero(y[i][0]).squeeze(-1)[0]
    work[i, :, :first_onset] = 0.
    outs.append(torchaudio.functional.resample(work[i, :, :cut], orig_freq=48000, new_freq=22050))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "postprocess_golden.npz"), gen=gen.numpy(), onsets=y.numpy(),
                    cut_length=cut, out=torch.stack(outs).numpy(), torchaudio=str(torchaudio.__version__))
print("ok", torch.stack(outs).shape)
