"""GPU-side input pipeline: raw uint16 tiles in HBM -> the fp32 [B, C, H, W] batch the encoder
consumes, without the per-image numpy / cv2 round trips of the reference.

Counterparts (same call shapes, same arithmetic; file:line of the reference):
  TileIngest        CARSThyroidDataset._preprocess_image   src/data/dataset.py:533-551      resize (cv2 INTER_LINEAR) + / 65535
                    AdaptiveNormalization('percentile')    src/data/quality_preprocessing.py:282-326
                    x.repeat(3,1,1) + T.Normalize           src/data/vit_transforms.py:381-393
  MixUp / CutMix    src/data/vit_transforms.py:396-462      host draws (np.random.beta, torch.randperm, np.random.randint)
                                                            are kept on the host exactly as in the reference; the
                                                            pixel work is one libvitk launch
The random PIL augmentations of `create_vit_transform` (flips, RandAugment, QualityAwarePatchAugment) are not part of this
module.  There is no CPU fallback: inputs must be CUDA tensors.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops

IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


class TileIngest:
    """raw uint16 [B, Hs, Ws] (CUDA) -> fp32 [B, channels, img_size, img_size].

    percentiles=None skips the adaptive normalisation; mean/std=None skips T.Normalize.  Defaults follow the pretrained
    transform branch of create_vit_transform (3 channels, ImageNet statistics)."""

    def __init__(self, img_size: int = 224, channels: int = 3, mean: Optional[Sequence[float]] = IMAGENET_MEAN,
                 std: Optional[Sequence[float]] = IMAGENET_STD, percentiles: Optional[Tuple[float, float]] = None):
        if img_size % 4:
            raise ValueError("img_size must be a multiple of 4")
        if (mean is None) != (std is None) or (mean is not None and (len(mean) != channels or len(std) != channels)):
            raise ValueError("mean / std: one value per output channel (or both None)")
        self.img_size, self.channels = int(img_size), int(channels)
        self.mean, self.std, self.percentiles = mean, std, percentiles

    def gray(self, raw: torch.Tensor) -> torch.Tensor:
        """_preprocess_image for the whole batch: fp32 [B, H, W] in [0, 1]."""
        if raw.dim() == 4 and raw.shape[1] == 1:
            raw = raw[:, 0]
        return ops.resize_u16(raw.contiguous(), self.img_size, self.img_size)

    def __call__(self, raw: torch.Tensor, mix: Optional[dict] = None) -> torch.Tensor:
        """mix (optional): {'perm': int tensor [B], 'lam': float} for MixUp or {'perm', 'box': (x1, y1, x2, y2)} for CutMix,
        applied after normalisation as the reference does on loader batches."""
        g = self.gray(raw)
        bounds = None
        if self.percentiles is not None:
            bounds = ops.percentile_bounds(g, self.percentiles[0] / 100, self.percentiles[1] / 100)
        kw = {}
        if mix is not None:
            kw["perm"] = mix["perm"].to(device=g.device, dtype=torch.int32)
            if "box" in mix:
                kw["cutmix"], kw["box"] = True, mix["box"]
            else:
                kw["lam"] = float(mix["lam"])
        return ops.finish_tiles(g, self.channels, bounds=bounds, mean=self.mean, std=self.std, **kw)


def _mix_planes(images: torch.Tensor, index: torch.Tensor, **kw) -> torch.Tensor:
    """Runs the mixing kernel over an already-built [B, C, H, W] batch: every (image, channel) plane is one 'tile'."""
    if not images.is_cuda:
        raise RuntimeError("MixUp / CutMix of this package run on CUDA tensors: there is no CPU fallback")
    B, C, H, W = images.shape
    planes = images.contiguous().float().view(B * C, H, W)
    perm = (index.to(images.device, torch.int64)[:, None] * C + torch.arange(C, device=images.device)[None, :]).reshape(-1)
    out = ops.finish_tiles(planes, 1, perm=perm.to(torch.int32), **kw)
    return out.view(B, C, H, W)


class MixUp:
    """vit_transforms.py:396-416: returns (mixed_images, labels_a, labels_b, lam)."""

    def __init__(self, alpha: float = 0.8):
        self.alpha = alpha

    def __call__(self, images: torch.Tensor, labels: torch.Tensor):
        batch_size = images.shape[0]
        lam = np.random.beta(self.alpha, self.alpha) if self.alpha > 0 else 1
        index = torch.randperm(batch_size)
        mixed = _mix_planes(images, index, lam=float(lam))
        return mixed, labels, labels[index.to(labels.device)], lam


class CutMix:
    """vit_transforms.py:419-462: returns (images, labels_a, labels_b, lam) with lam re-derived from the box area."""

    def __init__(self, alpha: float = 1.0):
        self.alpha = alpha

    def __call__(self, images: torch.Tensor, labels: torch.Tensor):
        batch_size = images.shape[0]
        lam = np.random.beta(self.alpha, self.alpha) if self.alpha > 0 else 1
        index = torch.randperm(batch_size)
        x1, y1, x2, y2 = self._rand_bbox(images.shape, lam)
        mixed = _mix_planes(images, index, cutmix=True, box=(x1, y1, x2, y2))
        lam = 1 - ((x2 - x1) * (y2 - y1) / (images.shape[-1] * images.shape[-2]))
        return mixed, labels, labels[index.to(labels.device)], lam

    def _rand_bbox(self, shape, lam):
        H, W = shape[2], shape[3]
        cut_rat = np.sqrt(1.0 - lam)
        cut_w, cut_h = np.int32(W * cut_rat), np.int32(H * cut_rat)
        cx, cy = np.random.randint(W), np.random.randint(H)
        return (int(np.clip(cx - cut_w // 2, 0, W)), int(np.clip(cy - cut_h // 2, 0, H)),
                int(np.clip(cx + cut_w // 2, 0, W)), int(np.clip(cy + cut_h // 2, 0, H)))
