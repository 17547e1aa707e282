"""f-3: a checkpoint whose keys follow another naming (the upstream closure-Module nesting, unknown
offline) is mapped onto the C-ABI names by registration order and shape - syncfusion_b200/checkpoint.py."""
import pytest
import torch

import syncfusion_b200 as sf
from syncfusion_b200.checkpoint import BLOCK_GROUPS, canonical_entries, structural_key_map
from syncfusion_b200.model import flat_param_name
from syncfusion_b200.synth import param_shapes, random_state_dict
from tests.util import SMALL, make_oracle


def _foreign(sd_flat, cfg, block_order, top_order, alias=True):
    """Re-key a flat state dict the way an unknown implementation might: opaque ``blocks.N`` names, another group
    order, the time MLP registered twice."""
    out = {}
    n = 0
    for name, shape, al in canonical_entries(cfg, block_order, top_order):
        if al is not None:
            if not alias:
                continue
            src = al
        else:
            src = name
        t = sd_flat[src]
        assert tuple(t.shape) == shape
        out[f"net.blocks.{n // 7}.blocks.{n % 7}.{'weight' if len(shape) > 1 else 'bias'}"] = t
        n += 1
    return out


@pytest.mark.parametrize("block_order", [BLOCK_GROUPS, ("skip", "down", "items_down", "inner", "items_up", "up"),
                                         ("down", "up", "skip", "items_down", "items_up", "inner")])
@pytest.mark.parametrize("top_order", [("unet", "time", "fixed"), ("time", "fixed", "unet"), ("fixed", "unet", "time")])
@pytest.mark.parametrize("alias", [True, False])
def test_structural_map_recovers_every_name(block_order, top_order, alias):
    cfg = sf.UNetConfig(**SMALL)
    flat = random_state_dict(cfg)
    foreign = _foreign(flat, cfg, block_order, top_order, alias)
    m = structural_key_map([(k, tuple(v.shape)) for k, v in foreign.items()], cfg, foreign)
    names = {v for v in m.values() if not v.endswith("#alias")}
    assert names == set(param_shapes(cfg))
    for k, name in m.items():
        assert torch.equal(foreign[k], flat[name.replace("#alias", "")])


def test_structural_map_full_architecture_and_oracle_order():
    cfg = sf.UNetConfig()                                     # exp/model/diffusion.yaml
    shapes = param_shapes(cfg)
    order = ("down", "items_down", "inner", "items_up", "up", "skip")
    keys = [(f"k{i}", s) for i, (n, s, al) in enumerate(canonical_entries(cfg, order)) if al is None]
    m = structural_key_map(keys, cfg)
    assert set(m.values()) == set(shapes) and len(m) == len(shapes)
    # the oracle's own state_dict order is one of the accepted registration orders
    om = make_oracle(SMALL)
    cfg_s = sf.UNetConfig(**SMALL)
    sd = om.net.state_dict()
    m = structural_key_map([(f"x.{i}", tuple(v.shape)) for i, v in enumerate(sd.values())], cfg_s)
    assert [m[f"x.{i}"] for i in range(len(sd))] == [flat_param_name(k) for k in sd]


def test_structural_map_failures_are_named():
    cfg = sf.UNetConfig(**SMALL)
    flat = random_state_dict(cfg)
    foreign = _foreign(flat, cfg, BLOCK_GROUPS, ("unet", "time", "fixed"))
    keys = [(k, tuple(v.shape)) for k, v in foreign.items()]
    with pytest.raises(KeyError, match="expected"):
        structural_key_map(keys[:40] + keys[41:], cfg)                      # one tensor missing
    cfg_t = sf.UNetConfig(upsample_mode="transpose", **SMALL)
    with pytest.raises(KeyError, match="upsample_mode='nearest'"):
        structural_key_map(keys, cfg_t)                                     # checkpoint of the other Upsample variant
    bad = dict(foreign)
    ks = [k for k, v in foreign.items() if tuple(v.shape) == (cfg.modulation_features, cfg.modulation_features)]
    bad[ks[1]] = foreign[ks[1]] + 1.0                                       # "alias" with different values
    with pytest.raises(KeyError, match="aliased"):
        structural_key_map(keys, cfg, bad)


def test_module_loads_a_foreign_checkpoint_through_the_parent():
    """main/generation.py:40-43 shape: Lightning-style parent, ``model.`` prefix, strict load."""
    cfg = sf.UNetConfig(**SMALL)
    flat = random_state_dict(cfg)
    foreign = _foreign(flat, cfg, ("skip", "down", "items_down", "inner", "items_up", "up"), ("time", "fixed", "unet"))

    class Parent(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model = sf.DiffusionModel(cfg)
            self.other = torch.nn.Linear(2, 2)

    parent = Parent()
    sd = {"model." + k: v for k, v in foreign.items()}
    sd.update({"other." + k: v for k, v in parent.other.state_dict().items()})
    res = parent.load_state_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys
    assert set(parent.model._staged) == set(param_shapes(cfg))
    for name, t in parent.model._staged.items():
        assert torch.equal(t, flat[name])
    assert parent.model.key_map["model.net.blocks.0.blocks.0.bias"] == "time.weights"       # first tensor of the first group
