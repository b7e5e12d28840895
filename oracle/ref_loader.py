"""TEST INFRASTRUCTURE -- loads the reference's own hand-written ViT/DeiT classes, unmodified,
from /root/reference (present only in the authoring container, never on the GPU box).

Used by oracle/make_golden.py (to generate tests/golden/*) and by tests/test_oracle.py (to pin
the restatement in oracle/vit_oracle.py against the real thing).  Nothing in the product package
imports this file.

Recipe: three stub packages (pytorch_lightning, torchmetrics, timm) on
sys.path, then the three source files are exec'd by path under their own module names with empty
parent packages so that src/models/__init__.py (which drags in the timm registry) is bypassed.
"""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

REF_ROOT = Path("/root/reference")
_STUBS = Path(__file__).resolve().parent / "_stubs"


def available() -> bool:
    return (REF_ROOT / "src/models/vit/deit_models.py").exists()


def load():
    """Returns (vision_transformer_base, vit_models, deit_models) modules of the reference."""
    if not available():
        raise RuntimeError("/root/reference is not mounted here")
    name0 = "src.models.vit.deit_models"
    if name0 in sys.modules and getattr(sys.modules[name0], "_oracle_loaded", False):
        return (sys.modules["src.models.vit.vision_transformer_base"], sys.modules["src.models.vit.vit_models"],
                sys.modules[name0])
    sys.dont_write_bytecode = True
    if str(_STUBS) not in sys.path:
        sys.path.insert(0, str(_STUBS))
    for pkg in ["src", "src.models", "src.models.vit"]:
        m = types.ModuleType(pkg)
        m.__path__ = [str(REF_ROOT / pkg.replace(".", "/"))]
        sys.modules[pkg] = m

    def _load(name, rel):
        spec = importlib.util.spec_from_file_location(name, REF_ROOT / rel)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    base = _load("src.models.vit.vision_transformer_base", "src/models/vit/vision_transformer_base.py")
    vitm = _load("src.models.vit.vit_models", "src/models/vit/vit_models.py")
    deit = _load("src.models.vit.deit_models", "src/models/vit/deit_models.py")
    deit._oracle_loaded = True
    return base, vitm, deit


class _RecordingAxes:
    """Stands in for a matplotlib Axes: keeps what the reference hands to imshow."""

    def __init__(self, sink):
        self._sink = sink

    def imshow(self, data, **kw):
        self._sink.append((data, kw))

    def set_title(self, *a, **k):
        pass

    def axis(self, *a, **k):
        pass


def load_attention_utils():
    """The reference's src/models/vit/attention_utils.py, unmodified.  matplotlib is absent here, so `matplotlib.pyplot`
    is replaced by a recorder: `subplots` returns axes whose `imshow` appends its argument to `module._imshow_calls`
    -- that is how the heat map `visualize_attention_maps` computes (attention_utils.py:50-67) is read back."""
    if not available():
        raise RuntimeError("/root/reference is not mounted here")
    name = "src.models.vit.attention_utils"
    if name in sys.modules and hasattr(sys.modules[name], "_imshow_calls"):
        return sys.modules[name]
    sys.dont_write_bytecode = True
    calls = []
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")

    class Figure:                       # the return annotation `-> plt.Figure` is evaluated at definition time
        pass

    def subplots(nrows=1, ncols=1, **kw):
        axes = [_RecordingAxes(calls) for _ in range(nrows * ncols)]
        return Figure(), (axes if nrows * ncols > 1 else axes[0])

    plt.Figure, plt.subplots = Figure, subplots
    plt.tight_layout = lambda *a, **k: None
    plt.savefig = lambda *a, **k: None
    mpl.pyplot = plt
    saved = {k: sys.modules.get(k) for k in ("matplotlib", "matplotlib.pyplot")}
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    try:
        spec = importlib.util.spec_from_file_location(name, REF_ROOT / "src/models/vit/attention_utils.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    mod._imshow_calls = calls
    sys.modules[name] = mod
    return mod


def reference_cls_heatmaps(attention_maps, image_hw, layer_idx=-1):
    """Runs the reference's visualize_attention_maps once per sample (it only ever looks at sample 0, :50) and returns
    the upsampled class-token maps it draws, [B, H_img, W_img] fp32."""
    import numpy as np
    import torch
    mod = load_attention_utils()
    out = []
    img = np.zeros(tuple(image_hw), dtype=np.float32)
    for b in range(attention_maps.shape[1]):
        del mod._imshow_calls[:]
        mod.visualize_attention_maps(attention_maps[:, b:b + 1], img, layer_indices=[layer_idx])
        heat = [d for d, kw in mod._imshow_calls if kw.get("cmap") == "hot"]
        assert len(heat) == 1
        out.append(torch.from_numpy(np.asarray(heat[0], dtype=np.float32)))
    return torch.stack(out)
