"""Generates tests/golden/small_unet_golden.pt from the oracle (the upstream packages cannot be imported here, so
these are self-generated pins: they freeze the oracle, they do not pin it against the reference).
    python tests/golden/make_golden.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.util import SMALL, make_inputs, make_oracle  # noqa: E402

if __name__ == "__main__":
    seed, B, L, steps, scale, sigma = 0, 1, 512, 3, 2.0, 0.6
    om = make_oracle(SMALL, stress=True, seed=seed)
    x, ch, e = make_inputs(om.net.cfg, B, L)
    with torch.no_grad():
        v = om.net(x, torch.full((B,), sigma), embedding=e, embedding_scale=scale, channels=ch)
        out = om.sample(x, num_steps=steps, channels=ch, embedding=e, embedding_scale=scale)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "small_unet_golden.pt")
    torch.save(dict(seed=seed, B=B, L=L, steps=steps, scale=scale, sigma=sigma, x=x, v=v, sample=out), path)
    print("wrote", path, os.path.getsize(path), "bytes")
