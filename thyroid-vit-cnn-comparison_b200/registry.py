"""ModelRegistry slot + wrapper classes -- the reference's drop-in boundary for models
(src/models/registry.py:19-98, src/models/base.py:9-51, src/models/vit/{__init__,deit,vision_transformer}.py).

`ModelRegistry` here has the reference's exact semantics (register decorator, create_model(config) ->
cls(config=config), ValueError for unknown names).  `install_into(ref_registry)` registers the B200
wrappers into the REFERENCE's own registry object (re-registration overwrites, registry.py:36-42), which is
all it takes to make ThyroidViTModule / ThyroidDistillationModule of the reference build B200 models.
"""
from __future__ import annotations

import logging
from typing import Any

import torch
import torch.nn as nn

from . import vit as V

logger = logging.getLogger(__name__)


class ModelRegistry:
    _registry: dict = {}

    @classmethod
    def register(cls, names, model_type: str = "default"):
        if not isinstance(names, list):
            names = [names]

        def decorator(model_class):
            cls._registry.setdefault(model_type, {})
            for name in names:
                cls._registry[model_type][name] = model_class
            return model_class
        return decorator

    @classmethod
    def create_model(cls, config):
        if not hasattr(config, "name"):
            raise ValueError("Configuration for model creation must include a 'name' attribute.")
        name = config.name
        for _, models in cls._registry.items():
            if name in models:
                return models[name](config=config)
        raise ValueError(f"Model '{name}' not found in registry. Available models: {cls.list_available_models()}")

    @classmethod
    def list_models(cls, model_type=None):
        if model_type:
            return list(cls._registry.get(model_type, {}).keys())
        return [n for models in cls._registry.values() for n in models]

    @classmethod
    def list_available_models(cls):
        return {t: list(m.keys()) for t, m in cls._registry.items()}


class ModelBase(nn.Module):
    """src/models/base.py:9-51."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.model = None

    def _build_model(self):
        raise NotImplementedError("Subclasses must implement _build_model.")

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.model is None:
            raise RuntimeError("The model has not been built yet. Ensure _build_model() is called in the subclass's "
                               "__init__ and assigns to self.model.")
        return self.model(x)


def _get(cfg: Any, key: str, default=None):
    if hasattr(cfg, "get"):
        try:
            v = cfg.get(key, default)
            return default if v is None else v
        except Exception:
            pass
    return getattr(cfg, key, default)


def _in_chans(cfg) -> int:
    """deit.py:29-32 / vision_transformer.py:31-35: extra_params['in_chans'], else config.channels, else 1 -- each tested
    with `is not None`, so an explicit value is never skipped."""
    extra = getattr(cfg, "extra_params", None) if hasattr(cfg, "extra_params") else None
    if extra is None and isinstance(cfg, dict):
        extra = cfg.get("extra_params")
    v = _get(extra, "in_chans", None) if extra is not None else None
    if v is None:
        v = _get(cfg, "channels", None)
    return 1 if v is None else int(v)


_TIMM_LN_EPS = 1e-6      # timm's VisionTransformer builds its norms as partial(nn.LayerNorm, eps=1e-6)


def _wrapper_kwargs(cfg) -> dict:
    """What the reference wrappers pass on (deit.py:26-53, vision_transformer.py:27-69): num_classes (default 2), img_size
    (224), patch_size (16, ViT wrapper only), in_chans.  The wrappers hand these to timm's plain ViT/DeiT variants --
    non-distilled, stochastic depth 0, LayerNorm eps 1e-6, no quality branch -- and never look at the YAML `params:` block,
    so the same holds here: the module tree (and therefore the state_dict key set a reference checkpoint was written
    with) is that of the timm model.

    `extra_params.handwritten: true` opts into the reference's HAND-WRITTEN constructors instead (deit_models.py /
    vit_models.py, which the YAML `params:` block -- embed_dim, depth, distilled, drop rates, quality_aware ... -- describes)."""
    import functools
    kw = dict(img_size=int(_get(cfg, "img_size", 224)), patch_size=int(_get(cfg, "patch_size", 16)),
              in_chans=_in_chans(cfg), num_classes=int(_get(cfg, "num_classes", 2)))
    extra = _get(cfg, "extra_params", {}) or {}
    if _get(extra, "handwritten", False):
        params = _get(cfg, "params", {}) or {}
        for k in ("embed_dim", "depth", "num_heads", "mlp_ratio", "qkv_bias", "drop_rate", "attn_drop_rate", "drop_path_rate",
                  "distilled", "quality_aware", "store_attention", "representation_size", "pos_embed_type", "pool_type",
                  "class_token"):
            v = _get(params, k, None)
            if v is not None:
                kw[k] = v
        return kw
    kw.update(drop_path_rate=0.0, quality_aware=False, norm_layer=functools.partial(nn.LayerNorm, eps=_TIMM_LN_EPS))
    return kw


@ModelRegistry.register(["deit_tiny", "deit_small", "deit_base"], "vit")
class DeiT(ModelBase):
    """src/models/vit/deit.py:10-72.  The reference maps these names to timm's NON-distilled deit_*_patch16_224 (one
    logits tensor; `lightning_modules.py:954-957` then uses it for both loss terms); see `_wrapper_kwargs`."""

    def __init__(self, config):
        super().__init__(config)
        self.variant = config.name
        self._build_model()

    def _build_model(self):
        name = self.config.name
        factory = {"deit_tiny": V.create_deit_tiny, "deit_small": V.create_deit_small, "deit_base": V.create_deit_base}.get(name)
        if factory is None:
            raise ValueError(f"Unsupported DeiT model name: {name}")
        kw = _wrapper_kwargs(self.config)
        kw.setdefault("distilled", False)
        kw.setdefault("drop_path_rate", 0.0)
        self.model = factory(pretrained=False, **kw)

    def get_parameter_groups(self, *a, **k):
        return self.model.get_parameter_groups(*a, **k)


@ModelRegistry.register(["vit_tiny", "vit_small", "vit_base"], "vit")
class VisionTransformer(ModelBase):
    """src/models/vit/vision_transformer.py:10-92."""

    def __init__(self, config):
        super().__init__(config)
        self.variant = config.name
        self._build_model()

    def _build_model(self):
        name = self.config.name
        if name not in V.VIT_MODEL_REGISTRY:
            raise ValueError(f"Unsupported ViT model name: {name}")
        kw = _wrapper_kwargs(self.config)
        kw.setdefault("drop_path_rate", 0.0)
        self.model = V.VIT_MODEL_REGISTRY[name](**kw)

    def get_parameter_groups(self, *a, **k):
        return self.model.get_parameter_groups(*a, **k)


def install_into(ref_registry) -> None:
    """Overwrite the reference registry's vit_* / deit_* entries with the B200 wrappers."""
    ref_registry.register(["deit_tiny", "deit_small", "deit_base"], "vit")(DeiT)
    ref_registry.register(["vit_tiny", "vit_small", "vit_base"], "vit")(VisionTransformer)


# --------------------------------------------------------------------------- checkpoint interop
_CKPT_PREFIXES = ("model.model.", "student.model.", "model.", "student.")


def lightning_state_dict(checkpoint, target: nn.Module) -> dict:
    """Key remapping for checkpoints written by the reference's Lightning modules.

    `ThyroidViTModule` stores the wrapper as `self.model` and the wrapper stores the network as `self.model`, so a `.ckpt`
    holds `model.model.<key>`; `scripts/run_ensemble_kfold_evaluation.py:98-101` rewrites that to `model.<key>` before
    `load_state_dict(strict=True)` on the wrapper.  `ThyroidDistillationModule` writes `student.model.<key>` (plus
    `teacher.*`, which is dropped).  `checkpoint` is the loaded dict (with or without the 'state_dict' level); the result
    is keyed for `target`, which may be a registry wrapper (keys `model.<key>`) or the bare network (keys `<key>`)."""
    sd = checkpoint.get("state_dict", checkpoint) if isinstance(checkpoint, dict) else checkpoint
    want = set(target.state_dict().keys())
    wrapper = isinstance(target, ModelBase)
    out = {}
    for key, value in sd.items():
        if key.startswith("teacher."):
            continue
        inner = key
        for pre in _CKPT_PREFIXES:
            if key.startswith(pre):
                inner = key[len(pre):]
                break
        new_key = ("model." + inner) if wrapper else inner
        if new_key in want:
            out[new_key] = value
        elif key in want:
            out[key] = value
    return out


def load_lightning_checkpoint(target: nn.Module, checkpoint, strict: bool = True):
    """load_model_for_fold's loading step (run_ensemble_kfold_evaluation.py:78-103) for an already-loaded checkpoint dict."""
    return target.load_state_dict(lightning_state_dict(checkpoint, target), strict=strict)
