"""Pins for the oracle: the reference holds no tests or golden vectors for this path and its
arithmetic lives in un-vendored packages, so parity is UNPINNED upstream; these closed-form and algebraic checks
plus the frozen fixtures in tests/golden/ are what pins the restatement."""
import math
import os

import pytest
import torch
import torch.nn as nn

from oracle import DiffusionModel, Encoder1d, UNetConfig, UNetV0, VSampler, count_parameters
from oracle.a_unet import Attention
from tests.trace import trace_unet
from tests.util import SMALL, make_inputs, make_oracle, rel_l2

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "small_unet_golden.pt")


class ZeroNet(nn.Module):
    def forward(self, x, t, **kw):
        return torch.zeros_like(x)


class ExactVNet(nn.Module):
    """v-prediction that points exactly at a fixed clean signal x*."""

    def __init__(self, target):
        super().__init__()
        self.target = target

    def forward(self, x, t, **kw):
        a = torch.cos(t * math.pi / 2).reshape(-1, 1, 1)
        b = torch.sin(t * math.pi / 2).reshape(-1, 1, 1)
        eps = (x - a * self.target) / b
        return a * eps - b * self.target


def test_sampler_zero_net_closed_form():
    x0 = torch.randn(3, 1, 64, dtype=torch.float64)
    for n in (1, 7, 50):
        out = VSampler(ZeroNet())(x0, n)
        assert torch.allclose(out, x0 * math.cos(math.pi / (2 * n)) ** n, rtol=1e-6, atol=1e-9)


def test_sampler_exact_v_returns_target():
    tgt = torch.randn(2, 1, 128, dtype=torch.float64)
    x0 = torch.randn(2, 1, 128, dtype=torch.float64)
    out = VSampler(ExactVNet(tgt))(x0, 10)
    assert rel_l2(out, tgt) < 1e-9


def test_sampler_trajectory_and_schedule():
    x0 = torch.randn(1, 1, 32)
    out, xs, vs = VSampler(ZeroNet())(x0, 4, return_trajectory=True)
    assert len(xs) == 5 and len(vs) == 4 and torch.equal(xs[-1], out)
    assert torch.equal(VSampler(ZeroNet()).schedule(4, "cpu"), torch.linspace(1.0, 0.0, 5))


def test_cfg_identities():
    om = make_oracle(SMALL, stress=True)
    x, ch, e = make_inputs(om.net.cfg, 2, 512)
    t = torch.full((2,), 0.4)
    v1 = om.net(x, t, embedding=e, embedding_scale=1.0, channels=ch)
    plain = om.net.xunet(x, om.net.time(t), e, ch)
    assert torch.equal(v1, plain)                                   # scale == 1: single pass, bit identical
    v0, v2, v3 = (om.net(x, t, embedding=e, embedding_scale=s, channels=ch) for s in (0.0, 2.0, 3.0))
    assert rel_l2(v3 - v2, v2 - v1) < 1e-4                          # affine in the scale
    mask = om.net.fixed_embedding.weight[None].expand(2, -1, -1)
    assert rel_l2(v0, om.net.xunet(x, om.net.time(t), mask, ch)) < 1e-5


def test_cross_attention_single_token_collapse():
    torch.manual_seed(0)
    att = Attention(32, 64, 8, context_features=512)
    x, e = torch.randn(2, 40, 32), torch.randn(2, 1, 512)
    full = att(x, e)
    v = att.to_kv(att.norm_ctx(e)).chunk(2, dim=-1)[1]
    assert rel_l2(full, x + att.to_out(v)) < 1e-6                   # softmax over one key == 1
    q = torch.randn(2, 8, 40, 64); k = torch.randn(2, 8, 5, 64)
    assert torch.allclose((torch.einsum("bhnd,bhmd->bhnm", q, k[:, :, :1])).softmax(-1), torch.ones(2, 8, 40, 1))


def test_attention_matches_sdpa():
    torch.manual_seed(1)
    att = Attention(64, 64, 8)
    x = torch.randn(2, 50, 64)
    xn, cn = att.norm(x), att.norm_ctx(x)
    q = att.to_q(xn); k, v = att.to_kv(cn).chunk(2, -1)
    q, k, v = (t.reshape(2, 50, 8, 64).transpose(1, 2) for t in (q, k, v))
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(2, 50, 512)
    assert rel_l2(att(x), x + att.to_out(o)) < 1e-5


def test_parameter_census():
    """exp/model/diffusion.yaml:16-33 -> 210,010,778 U-Net body parameters (transpose-upsample variant)."""
    net = UNetV0(UNetConfig(upsample_mode="transpose"))
    assert count_parameters(net.blocks) == 210_010_778
    assert count_parameters(net.time) == 128 + 257 * 1024 + 1024 + 1024 * 1024 + 1024   # tied MLP counted once
    assert count_parameters(net.fixed_embedding) == 512


def test_encoder_pyramid_matches_context_channels():
    """main/generation.py:80: xs[2:-1] must line up with exp/model/diffusion.yaml:22 context_channels."""
    enc = Encoder1d()
    _, info = enc(torch.zeros(1, 1, 2048), with_info=True)
    xs = info["xs"]
    assert len(xs) == 11
    got = [(t.shape[1], t.shape[2]) for t in xs[2:-1]]
    want = list(zip([2, 8, 16, 32, 64, 128, 256, 256], [2048 // f for f in (1, 4, 16, 64, 128, 256, 512, 1024)]))
    assert got == want


def test_inject_shape_assert_and_missing_embedding():
    om = make_oracle(SMALL)
    x, ch, e = make_inputs(om.net.cfg, 1, 256)
    t = torch.zeros(1)
    with pytest.raises(AssertionError):
        om.net(x, t, embedding=None, channels=ch)
    bad = list(ch); bad[1] = bad[1][:, :, :-1]
    with pytest.raises(AssertionError):
        om.net(x, t, embedding=e, channels=bad)


def test_trace_equals_forward():
    om = make_oracle(SMALL, stress=True)
    x, ch, e = make_inputs(om.net.cfg, 2, 512)
    t = torch.full((2,), 0.7)
    for s in (1.0, 2.0):
        tr, v = trace_unet(om.net, x, t, e, ch, s)
        assert rel_l2(v, om.net(x, t, embedding=e, embedding_scale=s, channels=ch)) < 1e-6


def test_fp64_floor():
    om = make_oracle(SMALL, stress=True)
    x, ch, e = make_inputs(om.net.cfg, 1, 512)
    t = torch.full((1,), 0.3)
    v32 = om.net(x, t, embedding=e, embedding_scale=2.0, channels=ch)
    om64 = make_oracle(SMALL, stress=True).double()
    v64 = om64.net(x.double(), t.double(), embedding=e.double(), embedding_scale=2.0, channels=[c.double() for c in ch])
    assert rel_l2(v32, v64) < 1e-4


def test_upsample_modes_differ_only_in_up():
    a = make_oracle(SMALL, upsample_mode="nearest").net.state_dict()
    b = make_oracle(SMALL, upsample_mode="transpose").net.state_dict()
    assert {k for k in a if ".up." not in k} == {k for k in b if ".up." not in k}


def test_golden_fixture():
    """Frozen outputs of this oracle (tests/golden/make_golden.py) - guards the restatement against drift."""
    g = torch.load(GOLDEN)
    om = make_oracle(SMALL, stress=True, seed=g["seed"])
    x, ch, e = make_inputs(om.net.cfg, g["B"], g["L"])
    assert torch.equal(x, g["x"])
    v = om.net(x, torch.full((g["B"],), g["sigma"]), embedding=e, embedding_scale=g["scale"], channels=ch)
    assert rel_l2(v, g["v"]) < 1e-5
    out = om.sample(x, num_steps=g["steps"], channels=ch, embedding=e, embedding_scale=g["scale"])
    assert rel_l2(out, g["sample"]) < 1e-5
