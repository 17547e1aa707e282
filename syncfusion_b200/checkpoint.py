"""Checkpoint ingestion for keys this package has never seen.

The reference loads ``torch.load(path)['state_dict']`` into the Lightning module (main/generation.py:40-43); the
``model.net.*`` entries of that dict are named by ``a_unet``'s closure-``Module`` nesting (``blocks.N.blocks.M...``),
whose exact spelling cannot be read offline (the package is not installed and the Zenodo checkpoint is not reachable).
What IS fixed by the architecture (exp/model/diffusion.yaml:11-33) is the *structure*: which
tensors exist, their shapes, and that ``state_dict()`` lists them in module-registration order, sub-module by
sub-module.  ``structural_key_map`` therefore maps a foreign key list onto the C-ABI names by order and shape alone:

* the U-Net is the recursive group ``{down, items_down, inner block, items_up, up, skip}`` and the wrappers add
  ``{time embedding, fixed (CFG) embedding}``; the registration ORDER of these groups is an upstream implementation
  detail, so every permutation is tried (the same one at every depth - it is one class) and a candidate survives only
  if all ~1000 shapes line up in sequence;
* the down-stack items are registered before the up-stack items (their shapes are identical, so this one ordering is
  assumed rather than searched);
* inside an item the order is the forward order (ResNet norm/conv pairs, Modulation linear, Inject conv, attention
  norm / norm_context / to_q / to_kv / to_out) - what a ``Sequential`` of the same blocks registers;
* the time MLP may appear twice (``Repeat`` registers the same ``Linear`` under two names): an aliased
  copy is accepted and checked for equality when the tensors are given.

The mapping is exact for any checkpoint whose tensors are laid out this way and fails loudly otherwise (first
mismatching position, both keys, both shapes).  It has been exercised against renamed / regrouped dicts of the oracle
(tests/test_checkpoint.py), not against the published checkpoint - that needs the file.
"""
from __future__ import annotations

from itertools import permutations
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

from torch import Tensor

from .config import UNetConfig

Shape = Tuple[int, ...]
Entry = Tuple[str, Shape, Optional[str]]     # (flat C-ABI name, shape, name this entry may alias)

BLOCK_GROUPS = ("down", "items_down", "inner", "items_up", "up", "skip")
TOP_GROUPS = ("unet", "time", "fixed")


def _item_entries(cfg: UNetConfig, d: int, q: str) -> List[Entry]:
    c, ctx, mf, ef = cfg.channels[d], cfg.context_channels[d], cfg.modulation_features, cfg.embedding_features
    mid = cfg.attention_heads * cfg.attention_features
    out: List[Entry] = []
    for n in ("1", "2"):
        out += [(q + f"resnet.gn{n}.weight", (c,), None), (q + f"resnet.gn{n}.bias", (c,), None),
                (q + f"resnet.conv{n}.weight", (c, c, 3), None), (q + f"resnet.conv{n}.bias", (c,), None)]
    out += [(q + "mod.linear.weight", (2 * c, mf), None), (q + "mod.linear.bias", (2 * c,), None)]
    if ctx > 0:
        out += [(q + "inject.conv.weight", (c, c + ctx, 1), None), (q + "inject.conv.bias", (c,), None)]
    kinds = ([("attn", c)] if cfg.attentions[d] else []) + ([("xattn", ef)] if cfg.cross_attentions[d] else [])
    for name, cf in kinds:
        a = q + f"{name}.attn."
        out += [(a + "norm.weight", (c,), None), (a + "norm.bias", (c,), None),
                (a + "norm_ctx.weight", (cf,), None), (a + "norm_ctx.bias", (cf,), None),
                (a + "to_q.weight", (mid, c), None), (a + "to_kv.weight", (2 * mid, cf), None),
                (a + "to_out.weight", (c, mid), None)]
    return out


def _block_entries(cfg: UNetConfig, d: int, order: Sequence[str], upsample_mode: str) -> List[Entry]:
    c, f, mf = cfg.channels[d], cfg.factors[d], cfg.modulation_features
    cin = cfg.in_channels if d == 0 else cfg.channels[d - 1]
    p = f"d{d}."
    groups: Dict[str, List[Entry]] = {
        "down": [(p + "down.weight", (c, cin, f), None), (p + "down.bias", (c,), None)],
        "up": ([(p + "up.weight", (c, cin, f), None), (p + "up.bias", (cin,), None)] if upsample_mode == "transpose" else
               [(p + "up.conv.weight", (cin, c, 3), None), (p + "up.conv.bias", (cin,), None)]),
        "skip": [(p + "skip.weight", (cin, mf), None), (p + "skip.bias", (cin,), None)],
        "inner": _block_entries(cfg, d + 1, order, upsample_mode) if d + 1 < cfg.depth else [],
    }
    for stack in ("items_down", "items_up"):
        groups[stack] = [e for i in range(cfg.items[d]) for e in _item_entries(cfg, d, f"{p}{stack}.{i}.")]
    return [e for g in order for e in groups[g]]


def canonical_entries(cfg: UNetConfig, block_order: Sequence[str] = BLOCK_GROUPS, top_order: Sequence[str] = TOP_GROUPS,
                      upsample_mode: Optional[str] = None) -> List[Entry]:
    """Every parameter of the architecture as (flat name, shape, alias) in one candidate registration order."""
    mf, ef = cfg.modulation_features, cfg.embedding_features
    top: Dict[str, List[Entry]] = {
        "unet": _block_entries(cfg, 0, block_order, upsample_mode or cfg.upsample_mode),
        "time": [("time.weights", (128,), None), ("time.linear.weight", (mf, 257), None), ("time.linear.bias", (mf,), None),
                 ("time.mlp.weight", (mf, mf), None), ("time.mlp.bias", (mf,), None),
                 ("time.mlp.weight#alias", (mf, mf), "time.mlp.weight"), ("time.mlp.bias#alias", (mf,), "time.mlp.bias")],
        "fixed": [("fixed_embedding.weight", (cfg.embedding_max_length, ef), None)],
    }
    return [e for g in top_order for e in top[g]]


def _block_orders():
    """Candidate registration orders of a block's groups.  The down and up item stacks have identical shapes, so shape
    alone cannot tell them apart: the down stack is taken to be registered before the up stack (both the forward order
    and the order of the constructor arguments)."""
    for order in permutations(BLOCK_GROUPS):
        if order.index("items_down") < order.index("items_up"):
            yield order


def _match(keys: Sequence[Tuple[str, Shape]], entries: Sequence[Entry]):
    """Sequential match; optional alias entries are consumed only if the next key has their shape.  Returns
    (mapping, None) or (None, (position, key, shape, wanted name, wanted shape))."""
    out: Dict[str, str] = {}
    i = 0
    for name, shape, alias in entries:
        if alias is not None:
            # an aliased copy directly follows the tensors it repeats; consume it only when the shapes fit AND skipping
            # it would not (the entry after the alias pair cannot have the alias shape in this architecture)
            if i < len(keys) and keys[i][1] == shape:
                out[keys[i][0]] = alias + "#alias"
                i += 1
            continue
        if i >= len(keys):
            return None, (i, "<end of checkpoint>", (), name, shape)
        if keys[i][1] != shape:
            return None, (i, keys[i][0], keys[i][1], name, shape)
        out[keys[i][0]] = name
        i += 1
    if i != len(keys):
        return None, (i, keys[i][0], keys[i][1], "<end of architecture>", ())
    return out, None


def structural_key_map(keys: Sequence[Tuple[str, Shape]], cfg: UNetConfig,
                       tensors: Optional[Mapping[str, Tensor]] = None) -> Dict[str, str]:
    """Map foreign ``state_dict`` keys (in ``state_dict()`` order, with shapes) onto flat C-ABI names by structure.

    Returns ``{checkpoint key: flat name}``; an aliased second copy of the time MLP maps to ``"<name>#alias"`` (callers
    drop those).  Raises ``KeyError`` with the best partial match if no registration order fits, or if two orders fit
    with different assignments (cannot happen for exp/model/diffusion.yaml: every group differs in rank or shape)."""
    keys = [(k, tuple(int(x) for x in s)) for k, s in keys]
    found: List[Tuple[Tuple[str, ...], Tuple[str, ...], Dict[str, str]]] = []
    best = (-1, None, None, None)
    for top in permutations(TOP_GROUPS):
        for order in _block_orders():
            m, miss = _match(keys, canonical_entries(cfg, order, top))
            if m is not None:
                if not any(m == f[2] for f in found):
                    found.append((top, order, m))
            elif miss[0] > best[0]:
                best = (miss[0], miss, top, order)
    if not found:
        other = "transpose" if cfg.upsample_mode == "nearest" else "nearest"
        hint = ""
        for top in permutations(TOP_GROUPS):
            for order in _block_orders():
                if _match(keys, canonical_entries(cfg, order, top, other))[0] is not None:
                    hint = f"; the checkpoint fits upsample_mode='{other}' - build the model with that mode"
                    break
            if hint:
                break
        pos, (i, k, s, want, wshape), top, order = best
        raise KeyError(f"checkpoint does not fit the architecture in any registration order{hint}: best candidate "
                       f"(groups {top} / {order}) matched {pos} of {len(keys)} tensors, then found '{k}' {s} where "
                       f"'{want}' {wshape} was expected")
    if len(found) > 1:
        canon = [f for f in found if f[0] == TOP_GROUPS and f[1] == BLOCK_GROUPS]
        if not canon:
            raise KeyError(f"{len(found)} registration orders fit the checkpoint with different assignments "
                           f"(e.g. {found[0][1]} and {found[1][1]}); pass flat names instead")
        found = canon
    mapping = found[0][2]
    if tensors is not None:                    # an alias must really be a copy
        inv = {v: k for k, v in mapping.items()}
        for k, v in mapping.items():
            if v.endswith("#alias"):
                src = inv[v[:-len("#alias")]]
                if not bool((tensors[k] == tensors[src]).all()):
                    raise KeyError(f"'{k}' has the shape of an aliased copy of '{src}' but different values")
    return mapping
