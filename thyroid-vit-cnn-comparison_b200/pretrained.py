"""Pretrained-weight interop for the drop-in DeiT -- what deit_models.py:109-188 does with a timm
ImageNet checkpoint, expressed as a small rule table instead of an if-chain, and fed from a LOCAL checkpoint (there is
no network and no timm on the deployment image; the reference downloads through `timm.create_model(pretrained=True)`).

Rules (reference line in brackets):
  * classifier tensors whose class count differs from the model's are dropped            [:147-149]
  * a position table of another length is resampled: prefix rows kept, patch rows taken as a sqrt x sqrt grid and
    resized bicubically (align_corners=False) to the model's grid                          [:152-155, :166-188]
  * RGB patch filters feeding a single-channel model are averaged over the colour axis     [:158-162]
Host-side, checkpoint-load time only; nothing here runs per step.
"""
from __future__ import annotations

import math
import warnings
from pathlib import Path
from typing import Dict, Mapping, Optional, Union

import torch
import torch.nn.functional as F

__all__ = ["resize_position_table", "adapt_state_dict", "read_checkpoint", "load_into"]


def resize_position_table(table: torch.Tensor, n_prefix: int, n_patches: int) -> torch.Tensor:
    """[1, n_prefix + g_old^2, D] -> [1, n_prefix + n_patches, D].  Identity when the patch counts already agree."""
    d = table.shape[-1]
    old = table.shape[1] - n_prefix
    if old == n_patches:
        return table
    g_old, g_new = int(math.sqrt(old)), int(math.sqrt(n_patches))
    grid = table[:, n_prefix:].reshape(1, g_old, g_old, d).permute(0, 3, 1, 2)          # raises if `old` is not a square
    grid = F.interpolate(grid, size=(g_new, g_new), mode="bicubic", align_corners=False)
    return torch.cat((table[:, :n_prefix], grid.permute(0, 2, 3, 1).reshape(1, g_new * g_new, d)), dim=1)


def adapt_state_dict(model, state_dict: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Apply the three rules above for `model` (a vit.DeiT / vit.VisionTransformer); key names are matched by substring,
    exactly as the reference does ('head' also catches 'head_dist')."""
    n_prefix = model._n_prefix()
    want_pos = tuple(model.pos_embed.shape)
    out: Dict[str, torch.Tensor] = {}
    for key, value in state_dict.items():
        if "head" in key and value.shape[0] != model.num_classes:
            continue
        if "pos_embed" in key and tuple(value.shape) != want_pos:
            value = resize_position_table(value, n_prefix, model.patch_embed.num_patches)
        if "patch_embed.proj.weight" in key and model.in_chans == 1 and value.dim() == 4 and value.shape[1] == 3:
            value = value.mean(dim=1, keepdim=True)
        out[key] = value
    return out


def read_checkpoint(source: Union[str, Path, Mapping[str, torch.Tensor]]) -> Mapping[str, torch.Tensor]:
    """A state_dict, or a file holding one -- bare, or under 'model' (the DeiT release files), 'state_dict' (Lightning)."""
    if isinstance(source, (str, Path)):
        blob = torch.load(str(source), map_location="cpu", weights_only=False)
        for k in ("model", "state_dict"):
            if isinstance(blob, dict) and k in blob and isinstance(blob[k], dict):
                return blob[k]
        return blob
    return source


def load_into(model, source: Optional[Union[str, Path, Mapping[str, torch.Tensor]]] = None):
    """DeiT.load_pretrained_weights (deit_models.py:109-139) with a local source.  Mirrors the reference's behaviour: no
    pretrained config and no source -> warn and skip; any failure -> warning, the model keeps its initialisation;
    otherwise load_state_dict(adapted, strict=False) and return the incompatible-keys record."""
    cfg = getattr(model, "pretrained_cfg", None) or {}
    if source is None:
        source = cfg.get("file") or cfg.get("checkpoint_path")
    if source is None:
        if not cfg:
            warnings.warn("No pretrained config provided, skipping weight loading")
        else:
            warnings.warn(f"pretrained weights for {cfg.get('model_name', '?')} cannot be downloaded here (no network / timm); "
                          "pass a local checkpoint: load_pretrained_weights(path_or_state_dict) or pretrained_cfg['file']")
        return None
    try:
        adapted = adapt_state_dict(model, read_checkpoint(source))
        result = model.load_state_dict(adapted, strict=False)
        if getattr(model, "_engine", None) is not None:
            model._sync_shadow()
        return result
    except Exception as e:  # noqa: BLE001 -- deit_models.py:138-139 turns every failure into a warning
        warnings.warn(f"Failed to load pretrained weights: {e}")
        return None
