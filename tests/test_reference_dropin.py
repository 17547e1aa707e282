"""Drop-in boundary against the UNMODIFIED reference sources (CPU; skipped where /root/reference does not exist).

The reference's own ``main.module_diffusion.Model`` (a Lightning module holding ``self.model``) and
``main.generation.generate_dataset`` are imported as they are; only packages that are not installable offline are
stubbed in ``sys.modules`` (pytorch_lightning, webdataset, librosa, plotly, omegaconf - none of them on the sampling
path) and ``audio_diffusion_pytorch.DiffusionModel`` / ``audio_encoders_pytorch.Encoder1d`` resolve to
``syncfusion_b200.DiffusionModel`` / the oracle encoder, which is exactly the swap INTEGRATION.md describes.

What runs for real: ``Model.__init__`` with our module as ``model``; ``Model.configure_optimizers`` (main/module_diffusion.py:53-62);
``model.load_state_dict(checkpoint['state_dict'])`` with Lightning-style ``model.net.*`` keys, strict (main/generation.py:40-43);
``model.to(device)`` (:44); the batch loop up to ``model.model.sample(...)`` (:77-83), whose keyword arguments and tensor
shapes are checked; and the post-processing / file naming after it (:85-122).  The CUDA call itself is replaced by a
recorder here (there is no GPU in this container; the same sequence runs on the GPU in tests/test_gpu_dropin.py).
"""
import importlib
import os
import sys
import types

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "main")), reason="reference sources not present")

from tests.util import SMALL, make_encoder, make_oracle  # noqa: E402


class _Anything:
    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return _Anything()
    def __getattr__(self, k): return _Anything()


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)

    def _missing(k):                        # any other public attribute: an inert placeholder class
        if k.startswith("__"):
            raise AttributeError(k)
        return _Anything

    m.__getattr__ = _missing
    sys.modules[name] = m
    return m


class _LightningModule(torch.nn.Module):
    """What generate_dataset needs of pl.LightningModule: an nn.Module with a ``device`` property that follows .to()."""

    def __init__(self):
        super().__init__()
        self.register_buffer("_pl_anchor", torch.empty(0), persistent=False)

    @property
    def device(self):
        return self._pl_anchor.device

    def log(self, *a, **k):
        pass


@pytest.fixture()
def reference(monkeypatch):
    import syncfusion_b200 as sf
    from oracle import Encoder1d
    names = ["pytorch_lightning", "pytorch_lightning.loggers", "pytorch_lightning.utilities", "webdataset", "webdataset.autodecode",
             "librosa", "plotly", "plotly.graph_objs", "omegaconf", "audio_diffusion_pytorch", "audio_encoders_pytorch"]
    saved = {k: sys.modules.get(k) for k in names}
    _stub("pytorch_lightning", LightningModule=_LightningModule, Callback=object, Trainer=object)
    _stub("pytorch_lightning.loggers", WandbLogger=type("WandbLogger", (), {}), Logger=object)
    _stub("pytorch_lightning.utilities", rank_zero_only=lambda f: f)
    _stub("webdataset", WebDataset=object)
    _stub("webdataset.autodecode", torch_audio=None)
    _stub("librosa")
    _stub("plotly")
    _stub("plotly.graph_objs")
    _stub("omegaconf", DictConfig=dict, OmegaConf=_Anything)
    _stub("audio_diffusion_pytorch", DiffusionModel=sf.DiffusionModel)
    _stub("audio_encoders_pytorch", Encoder1d=Encoder1d)
    monkeypatch.syspath_prepend(REF)
    for k in [k for k in sys.modules if k == "main" or k.startswith("main.")]:
        del sys.modules[k]
    gen = importlib.import_module("main.generation")
    mod = importlib.import_module("main.module_diffusion")
    yield gen, mod
    for k in [k for k in sys.modules if k == "main" or k.startswith("main.")]:      # only what this fixture put there
        del sys.modules[k]
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


class _Clap(torch.nn.Module):
    def load_ckpt(self, path):
        self.loaded = path

    def get_audio_embedding_from_data(self, x, use_tensor=True):
        g = torch.Generator().manual_seed(int(x.shape[-1]))
        e = torch.randn(x.shape[0], 512, generator=g)
        return e / e.norm(dim=-1, keepdim=True)

    def get_text_embedding(self, text, use_tensor=True):
        return self.get_audio_embedding_from_data(torch.zeros(len(text), 7))


def test_generate_dataset_runs_unmodified_against_the_shim(reference, tmp_path, monkeypatch):
    import syncfusion_b200 as sf
    import torchaudio
    gen, mod = reference
    L, B = 2048, 2
    cfg = sf.UNetConfig(precision="bf16", **SMALL)
    om = make_oracle(SMALL)
    enc = make_encoder(om.net.cfg)
    # the reference's own Lightning module around the B200 shim (exp/model/diffusion.yaml: model / onsets_encoder / embedder)
    model = mod.Model(lr=1e-4, lr_beta1=0.95, lr_beta2=0.999, lr_eps=1e-6, lr_weight_decay=1e-3, model=sf.DiffusionModel(cfg),
                      onsets_encoder=enc, embedder=_Clap(), embedder_checkpoint="clap.pt")
    assert isinstance(model.model, torch.nn.Module) and list(model.model.parameters()) == []
    opt = model.configure_optimizers()                                  # main/module_diffusion.py:53-62
    assert sum(p.numel() for g in opt.param_groups for p in g["params"]) == sum(p.numel() for p in enc.parameters())

    # a Lightning checkpoint: 'state_dict' with model.net.* (U-Net), onsets_encoder.* and no clap.* keys
    sd = {"model.net." + k: v for k, v in om.net.state_dict().items()}
    sd.update({"onsets_encoder." + k: v for k, v in enc.state_dict().items()})
    ckpt = tmp_path / "epoch=1.ckpt"
    torch.save({"state_dict": sd}, ckpt)

    calls = []

    def fake_sample(*, x_noisy, num_steps, channels, embedding, embedding_scale):      # stands in for the CUDA call only
        calls.append(dict(x=x_noisy, steps=num_steps, channels=channels, emb=embedding, scale=embedding_scale))
        return x_noisy * 0.5

    monkeypatch.setattr(model.model, "sample", fake_sample)
    saved_wavs = []
    monkeypatch.setattr(torchaudio, "save", lambda path, wav, sample_rate: saved_wavs.append((str(path), tuple(wav.shape), sample_rate)))

    g = torch.Generator().manual_seed(0)
    items = []
    for i in range(3):
        y = torch.zeros(1, L)
        y[0, torch.randint(100, L, (5,), generator=g)] = 1.0
        items.append((torch.randn(1, L, generator=g), y, torch.randn(1, 300 + 10 * i, generator=g), f"text{i}", f"dir/clip{i}"))
    gen.generate_dataset(tmp_path / "out", model, items, device="cpu", model_path=str(ckpt), batch_size=B, num_workers=0,
                         sample_rate=48000, num_steps=7, length=L, embedding_scale=2.0, cut_prefix=True, cut_length=1024,
                         downsample_rate=22050)
    # state dict reached the shim through nn.Module.load_state_dict of the PARENT, strictly, and was staged completely
    from syncfusion_b200.synth import param_shapes
    assert set(model.model._staged) == set(param_shapes(cfg))
    assert torch.equal(model.model._staged["d2.items_down.0.resnet.conv1.weight"], om.net.state_dict()["blocks.inner.inner.items_down.0.resnet.conv1.weight"])
    # two batches (2 + 1 clips), the keyword arguments of main/generation.py:77-83
    assert [c["x"].shape[0] for c in calls] == [2, 1]
    c0 = calls[0]
    assert c0["steps"] == 7 and c0["scale"] == 2.0 and tuple(c0["x"].shape) == (2, 1, L) and tuple(c0["emb"].shape) == (2, 1, 512)
    assert [tuple(t.shape) for t in c0["channels"]] == [(2, cfg.context_channels[d], cfg.length_at(L, d)) for d in range(cfg.depth)]
    assert len(saved_wavs) == 3 and saved_wavs[0][0].endswith("0.wav") and saved_wavs[0][2] == 22050


def test_strict_loading_reports_missing_and_unexpected_keys(reference):
    import syncfusion_b200 as sf
    _, mod = reference
    cfg = sf.UNetConfig(precision="bf16", **SMALL)
    om = make_oracle(SMALL)
    enc = make_encoder(om.net.cfg)
    model = mod.Model(lr=1e-4, lr_beta1=0.95, lr_beta2=0.999, lr_eps=1e-6, lr_weight_decay=1e-3, model=sf.DiffusionModel(cfg),
                      onsets_encoder=enc, embedder=_Clap(), embedder_checkpoint="x")
    sd = {"model.net." + k: v for k, v in om.net.state_dict().items()}
    sd.update({"onsets_encoder." + k: v for k, v in enc.state_dict().items()})
    bad = dict(sd)
    del bad["model.net.time.mlp.bias"]
    bad["model.net.blocks.items_down.0.resnet.convX.weight"] = torch.zeros(3)
    with pytest.raises(RuntimeError) as ei:
        model.load_state_dict(bad)
    assert "model.time.mlp.bias" in str(ei.value) and "convX" in str(ei.value)
    wrong = dict(sd)
    wrong["model.net.time.mlp.bias"] = torch.zeros(5)
    with pytest.raises(RuntimeError, match="size mismatch"):
        model.load_state_dict(wrong)
    res = model.load_state_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys
    with pytest.raises(Exception, match="CPU"):            # no CPU fallback: sampling on a CPU-resident module fails loudly
        model.model.sample(x_noisy=torch.zeros(1, 1, 512), num_steps=1, channels=[], embedding=torch.zeros(1, 1, 512), embedding_scale=1.0)
