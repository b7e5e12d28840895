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
