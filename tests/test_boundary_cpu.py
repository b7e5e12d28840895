"""Boundary tests that need no GPU: the registry wrappers against the reference's OWN ModelRegistry (loaded by path when
/root/reference is mounted), the stock-YAML wrapper defaults, and the pretrained-weight adapters against the reference's
own methods (oracle/ref_loader.py)."""
import importlib.util
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import thyroid_vit_cnn_comparison_b200  # noqa: E402,F401
from thyroid_vit_cnn_comparison_b200 import registry, vit, pretrained  # noqa: E402
from oracle import ref_loader  # noqa: E402

REF = Path("/root/reference")


class Cfg:                                                 # tests/unit/test_models.py:10-22 config contract
    def __init__(self, **k):
        self.__dict__.update(k)

    def get(self, k, d=None):
        return getattr(self, k, d)


def timm_vit_keys(depth: int):
    """state_dict key set of timm 1.0.15's VisionTransformer (what `timm.create_model('deit_tiny_patch16_224')` returns and
    what a reference checkpoint of the wired registry path therefore holds, under the wrapper's `model.` prefix)."""
    keys = ["cls_token", "pos_embed", "patch_embed.proj.weight", "patch_embed.proj.bias"]
    for i in range(depth):
        for sub in ("norm1", "attn.qkv", "attn.proj", "norm2", "mlp.fc1", "mlp.fc2"):
            keys += [f"blocks.{i}.{sub}.weight", f"blocks.{i}.{sub}.bias"]
    return keys + ["norm.weight", "norm.bias", "head.weight", "head.bias"]


def test_wrapper_defaults_follow_the_reference_wrappers_with_the_stock_yaml():
    """configs/model/vit/deit_tiny.yaml as shipped: name + a `params:` block (in_chans 1, distilled true, drop_path 0.1) the
    reference wrapper never reads (deit.py:26-53).  It builds timm's 1-channel, single-head, non-distilled model; so must we,
    and a checkpoint with that key set must load strictly."""
    import yaml
    y = REF / "configs/model/vit/deit_tiny.yaml"
    params = yaml.safe_load(y.read_text())["params"] if y.exists() else {"in_chans": 1, "distilled": True, "drop_path_rate": 0.1}
    w = registry.ModelRegistry.create_model(Cfg(name="deit_tiny", pretrained=False, params=params))
    net = w.model
    assert net.patch_embed.proj.in_channels == 1 and not hasattr(net, "dist_token") and not hasattr(net, "head_dist")
    assert tuple(net.pos_embed.shape) == (1, 197, 192) and net.norm.eps == 1e-6
    assert all(float(getattr(b.drop_path, "drop_prob", 0.0)) == 0.0 for b in net.blocks)
    assert sorted(w.state_dict().keys()) == sorted("model." + k for k in timm_vit_keys(12))     # same set (order is immaterial to load)
    g = torch.Generator().manual_seed(0)
    ckpt = {k: torch.randn(v.shape, generator=g) for k, v in w.state_dict().items()}
    res = w.load_state_dict(ckpt, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    # in_chans resolution order (deit.py:29-32): extra_params.in_chans, then channels, then 1 -- `is not None`, not truthiness
    assert registry._in_chans(Cfg(name="x", extra_params={"in_chans": 3}, channels=1)) == 3
    assert registry._in_chans(Cfg(name="x", extra_params={}, channels=3)) == 3
    assert registry._in_chans(Cfg(name="x")) == 1
    wv = registry.ModelRegistry.create_model(Cfg(name="vit_small", num_classes=10, img_size=224, patch_size=16, channels=3))
    assert wv.model.patch_embed.proj.in_channels == 3 and wv.model.head.out_features == 10
    assert sorted(wv.state_dict().keys()) == sorted("model." + k for k in timm_vit_keys(12))
    # opt-in: the hand-written constructors the YAML `params:` block describes
    wh = registry.ModelRegistry.create_model(Cfg(name="deit_tiny", extra_params={"handwritten": True, "in_chans": 1}, params=params))
    assert hasattr(wh.model, "dist_token") and tuple(wh.model.pos_embed.shape) == (1, 198, 192)
    assert wh.model.norm.eps == 1e-5 and any("quality_score" in k for k in wh.state_dict())
    assert wh.model.blocks[-1].drop_path.drop_prob == pytest.approx(0.1)


@pytest.mark.skipif(not (REF / "src/models/registry.py").exists(), reason="/root/reference not mounted")
def test_install_into_the_reference_registry_itself():
    """install_into() against the reference's real ModelRegistry object (src/models/registry.py imports stand-alone):
    re-registration overwrites (:36-42), create_model(config) then returns the B200 wrapper, unknown names still raise."""
    spec = importlib.util.spec_from_file_location("_ref_registry", REF / "src/models/registry.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    RefRegistry = mod.ModelRegistry

    @RefRegistry.register(["deit_tiny", "deit_small", "deit_base"], "vit")
    class Placeholder:                       # stands for the reference's timm-backed class already registered at import
        def __init__(self, config):
            raise AssertionError("the reference implementation must have been replaced")

    registry.install_into(RefRegistry)
    assert RefRegistry._registry["vit"]["deit_tiny"] is registry.DeiT
    assert RefRegistry._registry["vit"]["vit_base"] is registry.VisionTransformer
    w = RefRegistry.create_model(Cfg(name="deit_small", pretrained=False, num_classes=2, img_size=224))
    assert isinstance(w, registry.DeiT) and w.model.embed_dim == 384 and w.config.name == "deit_small"
    with pytest.raises(ValueError):
        RefRegistry.create_model(Cfg(name="no_such_model"))
    with pytest.raises(ValueError):
        RefRegistry.create_model(object())                     # no `name` attribute (:58-60)
    assert "deit_tiny" in RefRegistry.list_models("vit") if hasattr(RefRegistry, "list_models") else True


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_pretrained_adapters_match_the_reference_methods():
    """_adapt_pretrained_weights / _interpolate_pos_embed against the reference's own methods (deit_models.py:141-188) on a
    synthetic ImageNet-style checkpoint: 1000-class heads dropped, 224 -> 256 position table resampled, RGB -> gray filters."""
    _, _, ref_deit = ref_loader.load()
    kw = dict(img_size=256, patch_size=16, in_chans=1, num_classes=2, embed_dim=64, depth=1, num_heads=1, distilled=True)
    torch.manual_seed(0)
    ref = ref_deit.DeiT(**kw)
    # reference defect: `_adapt_pretrained_weights` reads `self.in_chans` (deit_models.py:158), which no reference class ever
    # sets (only hparams.in_chans exists), so as shipped it raises AttributeError on the first patch-filter key and
    # load_pretrained_weights degrades to its warning.  The attribute is supplied here to compare the intended arithmetic.
    assert not hasattr(ref, "in_chans")
    ref.in_chans = kw["in_chans"]
    ours = vit.DeiT(**kw)
    g = torch.Generator().manual_seed(1)
    ckpt = {
        "cls_token": torch.randn(1, 1, 64, generator=g), "dist_token": torch.randn(1, 1, 64, generator=g),
        "pos_embed": torch.randn(1, 2 + 14 * 14, 64, generator=g),                       # trained at 224
        "patch_embed.proj.weight": torch.randn(64, 3, 16, 16, generator=g), "patch_embed.proj.bias": torch.randn(64, generator=g),
        "blocks.0.attn.qkv.weight": torch.randn(192, 64, generator=g),
        "head.weight": torch.randn(1000, 64, generator=g), "head.bias": torch.randn(1000, generator=g),
        "head_dist.weight": torch.randn(1000, 64, generator=g), "head_dist.bias": torch.randn(1000, generator=g),
        "norm.weight": torch.randn(64, generator=g),
    }
    a, b = ref._adapt_pretrained_weights(dict(ckpt)), ours._adapt_pretrained_weights(dict(ckpt))
    assert list(a.keys()) == list(b.keys()) and "head.weight" not in b and "head_dist.bias" not in b
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert b["pos_embed"].shape == (1, 2 + 16 * 16, 64) and b["patch_embed.proj.weight"].shape == (64, 1, 16, 16)
    pe = torch.randn(1, 2 + 24 * 24, 64, generator=g)                                     # a 384-trained table
    assert torch.equal(ref._interpolate_pos_embed(pe), ours._interpolate_pos_embed(pe))
    same = torch.randn(1, 2 + 16 * 16, 64, generator=g)
    assert ours._interpolate_pos_embed(same) is same                                      # :171-172 identity
    # a 2-class checkpoint keeps its heads; a 3-channel model keeps RGB filters
    keep = ours._adapt_pretrained_weights({"head.weight": torch.zeros(2, 64), "head.bias": torch.zeros(2)})
    assert set(keep) == {"head.weight", "head.bias"}
    rgb = vit.DeiT(**{**kw, "in_chans": 3})
    assert rgb._adapt_pretrained_weights({"patch_embed.proj.weight": ckpt["patch_embed.proj.weight"]})["patch_embed.proj.weight"].shape[1] == 3


def test_load_pretrained_weights_from_a_local_checkpoint(tmp_path):
    kw = dict(img_size=64, patch_size=16, in_chans=1, num_classes=2, embed_dim=64, depth=1, num_heads=1, distilled=True)
    donor = vit.DeiT(**{**kw, "img_size": 32, "in_chans": 3, "num_classes": 5}, quality_aware=False)   # timm-style: no quality branch
    torch.save({"model": donor.state_dict()}, tmp_path / "donor.pth")
    m = vit.DeiT(**kw, pretrained_cfg={"model_name": "synthetic", "file": str(tmp_path / "donor.pth")})
    before_head = m.head.weight.detach().clone()
    res = m.load_pretrained_weights()
    assert res is not None and set(res.missing_keys) >= {"head.weight", "head.bias", "head_dist.weight", "head_dist.bias"}
    assert torch.equal(m.head.weight, before_head)                                        # 5-class heads were skipped
    assert torch.equal(m.blocks[0].attn.qkv.weight, donor.blocks[0].attn.qkv.weight)
    assert torch.allclose(m.patch_embed.proj.weight, donor.patch_embed.proj.weight.mean(1, keepdim=True))
    assert tuple(m.pos_embed.shape) == (1, 2 + 16, 64)
    with pytest.warns(UserWarning):                                                       # no config, no source: warn + skip
        assert vit.DeiT(**kw).load_pretrained_weights() is None
    with pytest.warns(UserWarning):                                                       # failures become warnings (:138-139)
        assert vit.DeiT(**kw).load_pretrained_weights(str(tmp_path / "missing.pth")) is None
