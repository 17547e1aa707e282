"""CPU/PyTorch oracle for SyncFusion's diffusion sampling hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
package, and only as the checker.  The product path (``syncfusion_b200``) never
imports it and fails loudly when its CUDA library is missing.

PARITY UNPINNED.  The reference repository (``/root/reference``) only configures
and calls the U-Net and the sampler (``exp/model/diffusion.yaml:11-33``,
``main/generation.py:77-83``, ``main/module_diffusion.py:200-206``); the arithmetic
lives in three pip packages that are neither vendored nor installable offline:

* ``audio-diffusion-pytorch==0.1.3``  (``requirements.txt:23``)  DiffusionModel, UNetV0, VDiffusion, VSampler, LinearSchedule
* ``a-unet`` (un-pinned transitive dependency; last published 0.0.16)  XUNet, Block, the *Item zoo, plugins
* ``audio-encoders-pytorch==0.0.22`` (``requirements.txt:24``)  Encoder1d

The reference holds no tests, golden vectors or fixtures for this path, so this
oracle restates the packages' published algorithm and is
pinned only by self-consistency tests (closed-form sampler answers, CFG identities,
the single-token cross-attention collapse, parameter census) and by the frozen
fixtures under ``tests/golden/`` that this very oracle generated.
"""
from .a_unet import UNetV0, UNetConfig, count_parameters, stress_init_  # noqa: F401
from .diffusion import DiffusionModel, VSampler, VDiffusion, LinearSchedule  # noqa: F401
from .encoder import Encoder1d  # noqa: F401
