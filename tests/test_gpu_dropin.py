"""f-1 and the whole drop-in chain on the GPU (need a B200).

* ``syncfusion_b200.Encoder1d`` (sfb_encoder_*: fused GroupNorm+SiLU+Conv1d streaming kernels) vs the oracle restatement of
  audio_encoders_pytorch.Encoder1d, every pyramid level, at the yaml's full configuration and L = 262144.
* The call sequence of main/generation.py:40-98 - Lightning-shaped parent module, ``load_state_dict`` of a checkpoint
  with ``model.net.*`` / ``onsets_encoder.*`` keys, ``.to(device)``, ``onsets_encoder(y, with_info=True)``,
  ``model.sample(channels=xs[2:-1], ...)``, mask / crop / resample - entirely on the GPU through this package, against
  the same sequence on the oracle.  (The unmodified reference file itself is driven in tests/test_reference_dropin.py,
  where /root/reference exists.)
Tolerances: encoder fp32 <= 1e-4 relative L2 per level (fp32 accumulation order); sampling as in test_gpu_parity.py."""
import numpy as np
import pytest
import torch

from oracle import Encoder1d as OracleEncoder
from oracle.postprocess import postprocess as oracle_postprocess
from tests.util import SMALL, make_encoder, make_inputs, make_oracle, rel_l2

pytestmark = pytest.mark.gpu


def _onsets(B, L, seed=0):
    g = torch.Generator().manual_seed(seed)
    y = torch.zeros(B, 1, L)
    for b in range(B):
        k = int(torch.randint(4, 17, (1,), generator=g))
        y[b, 0, torch.randint(0, min(L, 96000), (k,), generator=g)] = 1.0
    return y


@pytest.mark.parametrize("B,L", [(2, 262144), (3, 8192), (1, 4096 + 1024)])
def test_encoder_matches_oracle_full_config(cuda_device, B, L):
    import syncfusion_b200 as sf
    torch.manual_seed(1)
    oe = OracleEncoder().eval()                                   # exp/model/diffusion.yaml:35-43 defaults
    with torch.no_grad():
        for p in oe.parameters():                                 # stress: non-trivial GroupNorm affines and biases
            if p.ndim == 1:
                p.add_(0.3 * torch.randn_like(p))
    enc = sf.Encoder1d()
    res = enc.load_state_dict(oe.state_dict())
    assert not res.missing_keys and not res.unexpected_keys
    enc.to(cuda_device)
    y = _onsets(B, L) + 0.01 * torch.randn(B, 1, L, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        z_ref, info_ref = oe(y, with_info=True)
    z, info = enc(y.to(cuda_device), with_info=True)
    torch.cuda.synchronize()
    assert len(info["xs"]) == len(info_ref["xs"]) == 11
    for i, (a, b) in enumerate(zip(info["xs"], info_ref["xs"])):
        assert tuple(a.shape) == tuple(b.shape), (i, a.shape, b.shape)
        assert rel_l2(a.cpu(), b) < 1e-4, (i, rel_l2(a.cpu(), b))
    assert rel_l2(z.cpu(), z_ref) < 1e-4
    assert [t.shape[1] for t in info["xs"][2:-1]] == [2, 8, 16, 32, 64, 128, 256, 256]


class _Parent(torch.nn.Module):
    """Lightning-shaped parent (main/module_diffusion.py:20-49): .model, .onsets_encoder, clap_encode_audio, .device."""

    def __init__(self, model, onsets_encoder):
        super().__init__()
        self.model, self.onsets_encoder = model, onsets_encoder
        self.register_buffer("_a", torch.empty(0), persistent=False)

    @property
    def device(self):
        return self._a.device

    def clap_encode_audio(self, z):
        g = torch.Generator().manual_seed(int(z.shape[-1]))
        e = torch.randn(z.shape[0], 512, generator=g)
        return (e / e.norm(dim=-1, keepdim=True)).unsqueeze(1).to(z.device)


def _generate(model, y, z, noise, num_steps, scale, device, postprocess):
    """The body of generate_dataset's batch loop (main/generation.py:68-98), restated for the test."""
    y, z = y.to(device), z.to(device)
    _, y_latent = model.onsets_encoder(y, with_info=True)
    z_latent = model.clap_encode_audio(z)
    gen = model.model.sample(x_noisy=noise.to(device), num_steps=num_steps, channels=y_latent["xs"][2:-1],
                             embedding=z_latent.to(device), embedding_scale=scale)
    return postprocess(gen, y)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_generation_sequence_on_gpu_matches_oracle(cuda_device, precision):
    import syncfusion_b200 as sf
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B, L, steps, scale, cut = 2, 8192, 5, 2.0, 6000
    om = make_oracle(SMALL, stress=True)
    oenc = make_encoder(om.net.cfg)
    ckpt = {"model.net." + k: v for k, v in om.net.state_dict().items()}
    ckpt.update({"onsets_encoder." + k: v for k, v in oenc.state_dict().items()})
    ctx = list(SMALL["context_channels"])
    ours = _Parent(sf.DiffusionModel(sf.UNetConfig(precision=precision, **SMALL)),
                   sf.Encoder1d(in_channels=1, channels=2, multipliers=[1] + [c // 2 for c in ctx], factors=list(SMALL["factors"]),
                                num_blocks=[2] * len(ctx), resnet_groups=2, patch_size=1))
    res = ours.load_state_dict(ckpt)                                    # main/generation.py:42-43, strict
    assert not res.missing_keys and not res.unexpected_keys
    ours.to(cuda_device)                                                # :44
    ref = _Parent(om, oenc).to(cuda_device)
    y = _onsets(B, L, seed=3)
    z = torch.randn(B, 1, 777, generator=torch.Generator().manual_seed(4))
    noise = torch.randn(B, 1, L, generator=torch.Generator().manual_seed(5))
    out = _generate(ours, y, z, noise, steps, scale, cuda_device,
                    lambda gen, yy: sf.postprocess(gen, yy, cut_prefix=True, cut_length=cut, sample_rate=48000, downsample_rate=22050))
    with torch.no_grad():
        out_ref = _generate(ref, y, z, noise, steps, scale, cuda_device,
                            lambda gen, yy: torch.from_numpy(oracle_postprocess(gen.cpu().numpy(), yy.cpu().numpy(), cut_prefix=True, cut_length=cut,
                                                                                sample_rate=48000, downsample_rate=22050)))
    torch.cuda.synchronize()
    assert tuple(out.shape) == tuple(out_ref.shape) == (B, 1, int(np.ceil(cut * 147 / 320)))
    assert rel_l2(out.cpu(), out_ref) < (2e-3 if precision == "fp32" else 5e-2)


def test_foreign_checkpoint_names_sample_identically(cuda_device):
    """f-3: the same weights under opaque closure-Module style names, another group order and an aliased time MLP load
    through the parent's strict ``load_state_dict`` (structural key mapping) and sample bit-identically."""
    import syncfusion_b200 as sf
    from syncfusion_b200.checkpoint import canonical_entries
    from syncfusion_b200.model import flat_param_name
    cfg = sf.UNetConfig(precision="bf16", **SMALL)
    om = make_oracle(SMALL, stress=True)
    flat = {flat_param_name(k): v for k, v in om.net.state_dict().items()}
    foreign = {}
    for n, (name, shape, alias) in enumerate(canonical_entries(cfg, ("skip", "down", "items_down", "inner", "items_up", "up"), ("time", "fixed", "unet"))):
        foreign[f"model.net.blocks.{n // 5}.blocks.{n % 5}.p"] = flat[alias or name]

    class Parent(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model = sf.DiffusionModel(cfg)

    a, b = Parent(), Parent()
    res = a.load_state_dict(foreign)
    assert not res.missing_keys and not res.unexpected_keys
    b.load_state_dict({"model.net." + k: v for k, v in om.net.state_dict().items()})
    a.to(cuda_device); b.to(cuda_device)
    x, ch, e = make_inputs(om.net.cfg, 2, 2048)
    kw = dict(num_steps=3, channels=[c.to(cuda_device) for c in ch], embedding=e.to(cuda_device), embedding_scale=2.0)
    out_a = a.model.sample(x_noisy=x.to(cuda_device), **kw)
    out_b = b.model.sample(x_noisy=x.to(cuda_device), **kw)
    assert torch.equal(out_a, out_b)
