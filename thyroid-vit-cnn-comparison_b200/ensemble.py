"""Fold-sharded ensemble inference with attention rollout (BASELINE.json config 5, SURVEY.md section 8 a14 / e).

  EnsembleInference         evaluate_ensemble's hot loop (scripts/run_ensemble_kfold_evaluation.py:127-152):
                            probs_m = softmax(model_m(images)); sum_m w_m * probs_m; argmax -- with the member models
                            sharded `fold f -> rank f mod world` and one tiny all-gather of logits as the only exchange
  EnsembleTeacher           src/utils/models.py:231-283: normalised weights, weighted sum of LOGITS (distillation teacher)
  create_attention_rollout  src/models/vit/attention_utils.py:129-145 (the reference body is `pass`; spec = Abnar & Zuidema 2020)
  cls_attention_grid        visualize_attention_maps' CLS-row map (attention_utils.py:49-62) without the matplotlib part
  cls_attention_heatmap     the same map upsampled to the image (attention_utils.py:50-67: mean over heads, CLS row, sqrt grid,
                            bilinear F.interpolate) for every image of the batch, one `vitk_cls_attention_heatmap` launch

Every forward is the libvitk eval path of the member model (`engine.forward(train=False)`); the probability mix and
the rollout are `vitk_ensemble_probs` / `vitk_attention_rollout`; torch only averages DeiT's two [B,classes] heads
(deit_models.py:233-238, as the model class itself does) and stacks the F per-fold results.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops
from .parallel import gather_fold_logits, shard_folds

__all__ = ["EnsembleInference", "EnsembleTeacher", "create_attention_rollout", "cls_attention_grid", "cls_attention_heatmap"]


def _inner(model: nn.Module) -> nn.Module:
    """Registry wrappers (registry.DeiT / VisionTransformer) keep the network in `.model` (src/models/base.py:9-51)."""
    return model if hasattr(model, "_ensure_engine") else getattr(model, "model")


def create_attention_rollout(attention_maps: torch.Tensor, head_fusion: str = "mean") -> torch.Tensor:
    """attention_maps fp32 [L,B,H,N,N] (what `get_attention_maps()` returns, vision_transformer_base.py:488-492) ->
    rollout [B,N,N]: fuse heads (mean / max / min), A_hat = (A + I)/2 row-normalised, R = A_hat_L ... A_hat_1."""
    if head_fusion not in ("mean", "max", "min"):
        raise ValueError(f"Unknown head_fusion: {head_fusion}")
    if attention_maps.dim() != 5 or attention_maps.shape[-1] != attention_maps.shape[-2]:
        raise ValueError("attention_maps must be [layers, batch, heads, N, N]")
    if not attention_maps.is_cuda:
        raise RuntimeError("create_attention_rollout runs on a CUDA device through libvitk.so (no CPU fallback)")
    return ops.attention_rollout(attention_maps.float().contiguous(), head_fusion)


def cls_attention_grid(rollout_or_map: torch.Tensor, n_prefix: int) -> torch.Tensor:
    """Row 0 (the class token) over the patch columns of a [B,N,N] matrix, as the sqrt-grid [B,g,g]."""
    a = rollout_or_map[:, 0, n_prefix:]
    g = int(math.isqrt(a.shape[-1]))
    if g * g != a.shape[-1]:
        raise ValueError("patch count is not a square grid")
    return a.reshape(-1, g, g)


def cls_attention_heatmap(attention: torch.Tensor, image_size, layer_idx: int = -1, n_prefix: int = 1, grid: bool = False) -> torch.Tensor:
    """The numeric part of `visualize_attention_maps` (attention_utils.py:50-67) for the whole batch -> fp32 [B,H_img,W_img].

    attention   [L,B,H,N,N] (what `get_attention_maps()` returns; `layer_idx` picks the layer, default the last one like
                the reference's `layer_indices=[-1]`), one layer's [B,H,N,N], a rollout matrix [B,N,N] (its class-token
                row is used) or a rollout row [B,N]; with grid=True a ready [B,g,g] grid, e.g. one fold of what
                `EnsembleInference(rollout=True)` returns (a 3-D input is ambiguous otherwise: [B,10,10] could be either).
    image_size  `original_image.shape[:2]` (int or (h, w)).
    n_prefix    tokens before the patches: the reference slices `attn[0, 1:]`, i.e. 1 -- with a distilled DeiT's 198 tokens
                that leaves 197 columns and its `reshape(14, 14)` raises; pass 2 there.  Ignored with grid=True.
    """
    if not attention.is_cuda:
        raise RuntimeError("cls_attention_heatmap runs on a CUDA device through libvitk.so (no CPU fallback)")
    hw = (int(image_size), int(image_size)) if isinstance(image_size, int) else (int(image_size[0]), int(image_size[1]))
    a = attention.float()
    if grid:
        if a.dim() != 3 or a.shape[-1] != a.shape[-2]:
            raise ValueError("grid=True expects [B,g,g]")
        return ops.cls_attention_heatmap(a, hw, 0)
    if a.dim() == 5:
        a = a[layer_idx]
    elif a.dim() == 3:
        if a.shape[-1] != a.shape[-2]:
            raise ValueError("a 3-D input must be a rollout matrix [B,N,N] (or a [B,g,g] grid with grid=True)")
        a = a[:, 0, :]                               # the class token's row
    elif a.dim() not in (2, 4):
        raise ValueError("attention must be [L,B,H,N,N], [B,H,N,N], [B,N,N] or [B,N]")
    if a.stride(-1) != 1:
        a = a.contiguous()
    return ops.cls_attention_heatmap(a, hw, n_prefix)


class EnsembleInference:
    """F member models, evaluated fold-sharded.

    models      the member models THIS rank owns, in the order of `shard_folds(num_folds, rank, world)`
                (single process: all F models).  Drop-in classes of this package (or their registry wrappers).
    weights     F ensemble weights (`evaluate_ensemble`'s `weights` tensor); default uniform 1/F (config 5).
    num_folds   F (default: len(models) * world for an even split, else required).
    rollout     also return the per-fold CLS attention-rollout grids [F,B,g,g].

    __call__(images[B,C,H,W]) -> dict(probs [B,classes], preds [B] int64, logits [F,B,classes],
                                      rollout [F,B,g,g] (optional)) -- identical on every rank.
    """

    def __init__(self, models: Sequence[nn.Module], weights: Optional[Sequence[float]] = None, num_folds: Optional[int] = None,
                 process_group=None, rollout: bool = False, head_fusion: str = "mean"):
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.models = [m.eval() for m in models]
        if num_folds is None:
            if self.world != 1:
                raise ValueError("num_folds is required when the folds are sharded across ranks")
            num_folds = len(self.models)
        self.num_folds = int(num_folds)
        self.local_folds = shard_folds(self.num_folds, self.rank, self.world)
        if len(self.local_folds) != len(self.models):
            raise ValueError(f"rank {self.rank} of {self.world} owns folds {self.local_folds} but got {len(self.models)} models")
        if weights is None:
            weights = [1.0 / self.num_folds] * self.num_folds
        if len(weights) != self.num_folds:
            raise AssertionError("Number of models must match number of weights.")     # run_ensemble_kfold_evaluation.py:136
        self._weights_host = torch.tensor([float(w) for w in weights], dtype=torch.float32)
        self._weights = None
        if head_fusion not in ("mean", "max", "min"):
            raise ValueError(f"Unknown head_fusion: {head_fusion}")
        self.rollout, self.head_fusion = rollout, head_fusion
        self._maps: Dict[tuple, torch.Tensor] = {}
        # ranks that own no fold (more ranks than folds) still need the gather shapes: rank 0 (always owns fold 0) tells them,
        # once, collectively, at construction time
        shape = [0, 0]
        if self.models:
            net = _inner(self.models[0])
            shape = [int(net.num_classes), int(math.isqrt(net.num_patches))]
        if self.world > 1:
            on_gpu = dist.get_backend(self.pg) == "nccl"
            t = torch.tensor(shape, dtype=torch.int64, device=torch.device("cuda", torch.cuda.current_device()) if on_gpu else "cpu")
            dist.broadcast(t, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)
            shape = [int(v) for v in t.tolist()]
        self._num_classes, self._grid = shape

    def _member_forward(self, model: nn.Module, images: torch.Tensor, want_maps: bool, gray=None):
        net = _inner(model)
        eng = net._ensure_engine()
        net._sync_shadow()
        maps = None
        if want_maps:
            d = eng.d
            key = (images.shape[0], d.depth, d.heads, d.tokens)
            maps = self._maps.get(key)
            if maps is None:                       # one image-major [B,L,H,T,T] buffer, reused by every member of that shape:
                # the rollout walks one image's L*H maps per CTA, so they are kept contiguous
                maps = torch.empty(images.shape[0], d.depth, d.heads, d.tokens, d.tokens, dtype=torch.float32, device=images.device)
                self._maps[key] = maps
        l0, l1 = eng.forward(images, train=False, attn_probs=None if maps is None else ("image_major", maps), gray=gray)
        if l1 is not None:                         # DeiT eval: mean of the cls and dist heads (deit_models.py:233-238)
            l0 = (l0 + l1) / 2
        return l0, maps, eng.d.n_prefix

    @torch.no_grad()
    def __call__(self, images: torch.Tensor, gray=None) -> Dict[str, torch.Tensor]:
        """gray (engine.GraySpec, optional): `images` are single-channel tiles [B,H,W] (raw uint16 / fp16 / fp32) that every
        member replicates + normalises on the device while writing its patch matrix (see VitEngine.forward)."""
        if not images.is_cuda:
            raise RuntimeError("EnsembleInference runs on a CUDA device through libvitk.so (no CPU fallback)")
        dev = images.device
        if self._weights is None or self._weights.device != dev:
            self._weights = self._weights_host.to(dev)
        logits, grids = [], []
        for m in self.models:
            lg, maps, n_prefix = self._member_forward(m, images, self.rollout, gray)
            logits.append(lg.float())
            if self.rollout:
                row = ops.attention_rollout_row(maps, 0, self.head_fusion, image_major=True)   # class-token row of the rollout, [B,N]
                g = int(math.isqrt(row.shape[1] - n_prefix))
                grids.append(row[:, n_prefix:].reshape(-1, g, g))
        B = images.shape[0]
        if logits:
            local = torch.stack(logits, dim=0)
        else:                                       # more ranks than folds: this rank only takes part in the gather
            local = torch.zeros(0, B, self._num_classes, dtype=torch.float32, device=dev)
        full = gather_fold_logits(local, self.local_folds, self.num_folds, self.pg).contiguous()
        probs, preds = ops.ensemble_probs(full, self._weights)
        out = {"probs": probs, "preds": preds, "logits": full}
        if self.rollout:
            g = grids[0].shape[-1] if grids else self._grid
            loc = torch.stack(grids, dim=0) if grids else torch.zeros(0, B, g, g, dtype=torch.float32, device=dev)
            out["rollout"] = gather_fold_logits(loc.reshape(loc.shape[0], B, g * g), self.local_folds, self.num_folds,
                                                self.pg).reshape(self.num_folds, B, g, g)
        return out

class EnsembleTeacher(nn.Module):
    """src/utils/models.py:231-283 -- weighted average of teacher LOGITS (weights normalised to sum 1)."""

    def __init__(self, teachers: List[nn.Module], weights: Optional[List[float]] = None, device: Optional[str] = None):
        super().__init__()
        self.teachers = nn.ModuleList(teachers)
        self.num_teachers = len(teachers)
        if weights is None:
            weights = [1.0 / self.num_teachers] * self.num_teachers
        else:
            total = sum(weights)
            weights = [w / total for w in weights]
        self.weights = torch.tensor(weights, device=device)

    def get_individual_predictions(self, x: torch.Tensor) -> List[torch.Tensor]:
        with torch.no_grad():
            return [t(x) for t in self.teachers]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        stacked = torch.stack([p.float() for p in self.get_individual_predictions(x)], dim=0)     # [n_teachers, B, classes]
        return (stacked * self.weights.to(stacked).view(-1, 1, 1)).sum(dim=0)
