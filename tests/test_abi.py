"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares,
the host wrapper maps parameter names, and the product path fails loudly without a GPU (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

import syncfusion_b200 as sf
from syncfusion_b200 import _lib
from syncfusion_b200.model import flat_param_name

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "syncfusion_b200.h")).read()
    return sorted(set(re.findall(r"\b(sfb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    from syncfusion_b200 import build
    lib_path = build.build()
    lib = ctypes.CDLL(str(lib_path))
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/syncfusion_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == syms            # the ctypes binding covers exactly the header


def test_config_struct_matches_header_layout():
    c = _lib.SfbUnetConfig()
    assert ctypes.sizeof(c) == 4 * (2 + 6 * 16 + 8)


def test_flat_param_names():
    assert flat_param_name("blocks.down.weight") == "d0.down.weight"
    assert flat_param_name("blocks.inner.inner.items_up.1.resnet.conv2.bias") == "d2.items_up.1.resnet.conv2.bias"
    assert flat_param_name("model.net.blocks.inner.skip.weight") == "d1.skip.weight"
    assert flat_param_name("net.time.mlp.weight") == "time.mlp.weight"
    assert flat_param_name("d3.up.conv.weight") == "d3.up.conv.weight"


def test_random_state_dict_matches_oracle_names_and_shapes():
    from oracle import UNetConfig as OC, UNetV0
    from tests.util import SMALL
    for mode in ("nearest", "transpose"):
        o = {flat_param_name(k): tuple(v.shape) for k, v in UNetV0(OC(upsample_mode=mode, **SMALL)).state_dict().items()}
        s = {k: tuple(v.shape) for k, v in sf.random_state_dict(sf.UNetConfig(upsample_mode=mode, **SMALL)).items()}
        assert o == s


def test_synthetic_inputs_shapes():
    cfg = sf.UNetConfig()
    x, ch, e = sf.synthetic_inputs(cfg, 2, 4096)
    assert x.shape == (2, 1, 4096) and e.shape == (2, 1, 512)
    assert [tuple(c.shape) for c in ch] == [(2, c, 4096 // f) for c, f in
                                            zip(cfg.context_channels, (1, 4, 16, 64, 128, 256, 512, 1024))]
    assert torch.allclose(e.norm(dim=-1), torch.ones(2, 1))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(_lib.SfbError):
        sf.DiffusionModel(sf.UNetConfig(), "cuda")
    lib = _lib.load()
    h = ctypes.c_void_p()
    cfg = sf.model._c_config(sf.UNetConfig())
    assert lib.sfb_create(ctypes.byref(cfg), 0, ctypes.byref(h)) != 0   # no device -> error status, never a CPU path


def test_product_code_never_imports_oracle():
    pkg = os.path.join(ROOT, "syncfusion_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn


def test_hydra_config_accepts_the_reference_yaml_keys():
    """exp/model/diffusion.yaml:11-33 keyword set -> UNetConfig (the Hydra `_target_` swap of INTEGRATION.md)."""
    import syncfusion_b200 as sf
    cfg = sf.hydra_config(net_t=object(), diffusion_t=object(), sampler_t=object(), in_channels=1,
                          channels=[8, 32, 64, 128, 256, 512, 1024, 1024], factors=[1, 4, 4, 4, 2, 2, 2, 2],
                          items=[1, 2, 2, 2, 2, 2, 2, 4], attentions=[0, 0, 0, 0, 1, 1, 1, 1], attention_heads=8,
                          attention_features=64, context_channels=[2, 8, 16, 32, 64, 128, 256, 256],
                          use_embedding_cfg=True, embedding_max_length=1, embedding_features=512,
                          cross_attentions=[1] * 8)
    assert cfg == sf.UNetConfig()
    with pytest.raises(TypeError):
        sf.hydra_config(unknown_key=1)


def test_library_sass_has_tcgen05_tma_and_no_other_arch():
    """The shipped .so is sm_100a-only machine code with tcgen05 MMAs (UTCHMMA), TMA loads / stores (UTMALDG / UTMASTG) and
    TMEM loads (LDTM) - the mnemonics /opt/skills/guides/B200_PROFILING.md names as proof of the native path."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    lib = str(_lib.LIB_PATH)
    elf = subprocess.run([cuobjdump, "-lelf", lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", elf))
    assert archs == {"100a"}, archs
    sass = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True).stdout      # ~10 s
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):
        assert mnemonic in sass, mnemonic

