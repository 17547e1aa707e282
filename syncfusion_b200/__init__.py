"""syncfusion_b200 - B200-native (sm_100a) drop-in for SyncFusion's diffusion sampling path.

Host-side mirror of the reference interface (``DiffusionModel.sample`` as called at
/root/reference/main/generation.py:77-83 and main/module_diffusion.py:200-206) over the C ABI of
``libsyncfusion_b200.so``.  PyTorch is used for device memory, streams and ``torch.distributed`` only.
"""
from .config import UNetConfig  # noqa: F401
from .model import DiffusionModel, UNetV0, VSampler, flat_param_name, hydra_config, hydra_target  # noqa: F401
from .synth import random_state_dict, synthetic_inputs  # noqa: F401
from .parallel import shard_batch, gather_waveforms, sample_sharded  # noqa: F401
from .postprocess import postprocess  # noqa: F401
from .encoder import Encoder1d  # noqa: F401

__all__ = ["UNetConfig", "DiffusionModel", "UNetV0", "VSampler", "flat_param_name", "shard_batch",
           "gather_waveforms", "sample_sharded", "hydra_config", "hydra_target", "random_state_dict", "synthetic_inputs", "postprocess", "Encoder1d"]
