"""On-device classification metrics.

Counterpart of the torchmetrics objects the reference's Lightning modules create in `_setup_metrics`
(lightning_modules.py:352-374, :895-920: Accuracy, AUROC, F1Score, Specificity, Recall, Precision, StatScores) and update
on the host in every validation / test step (:491-516, :542-560, :1003-1030).  Here one tiny libvitk launch per step
updates integer counters that live on the GPU (`vitk_metrics_update`); nothing is copied to the host until `compute()`,
which launches the pairwise AUROC count (`vitk_binary_auroc`) and reads back a handful of integers.

Definitions follow torchmetrics 1.7.2 (requirements.txt:171) for task='binary' with hard predictions:
  acc = (tp+tn)/(tp+tn+fp+fn)   f1 = 2tp/(2tp+fp+fn)   specificity = tn/(tn+fp)   sensitivity (Recall) = tp/(tp+fn)
  ppv (Precision) = tp/(tp+fp)  -- every ratio is 0 when its denominator is 0 (`_safe_divide`)
  stat_scores = [tp, fp, tn, fn, support = tp+fn]            npv = tn/(tn+fn+1e-6)   (hand-written at :512-515)
  auc = area under the exact ROC curve of softmax(logits)[:, 1]; 0 when a class is absent.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops


def _safe_divide(num: float, den: float) -> float:
    return num / den if den != 0 else 0.0


def binary_metrics_from_counts(tp: int, fp: int, tn: int, fn: int) -> Dict[str, float]:
    return {
        "acc": _safe_divide(tp + tn, tp + tn + fp + fn),
        "f1": _safe_divide(2 * tp, 2 * tp + fp + fn),
        "specificity": _safe_divide(tn, tn + fp),
        "sensitivity": _safe_divide(tp, tp + fn),
        "ppv": _safe_divide(tp, tp + fp),
        "npv": tn / (tn + fn + 1e-6),
        "stat_scores": [tp, fp, tn, fn, tp + fn],
    }


class ClassificationMetrics:
    """Accumulates one split's metrics on the device.  `capacity` bounds the number of samples kept for the AUROC
    (binary task only); exceeding it raises at compute() rather than silently truncating."""

    def __init__(self, num_classes: int = 2, capacity: int = 1 << 16, device=None):
        if num_classes < 2:
            raise ValueError("num_classes must be >= 2")
        self.num_classes = int(num_classes)
        self.capacity = int(capacity)
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if dev.type != "cuda":
            raise RuntimeError("ClassificationMetrics lives on a CUDA device: there is no CPU fallback")
        self.confusion = torch.zeros(self.num_classes ** 2 + 1, dtype=torch.int64, device=dev)
        self.binary = self.num_classes == 2
        if self.binary:
            self.scores = torch.empty(self.capacity, dtype=torch.float32, device=dev)
            self.score_labels = torch.empty(self.capacity, dtype=torch.uint8, device=dev)
            self.count = torch.zeros(1, dtype=torch.int64, device=dev)
            self._scratch = torch.empty(4, dtype=torch.int64, device=dev)
            self._auc = torch.empty(4, dtype=torch.float64, device=dev)

    def reset(self) -> None:
        self.confusion.zero_()
        if self.binary:
            self.count.zero_()

    def update(self, logits: torch.Tensor, labels: torch.Tensor) -> None:
        """logits [B, C] (any float dtype, CUDA), labels [B] / [B,1] integer."""
        if logits.dim() != 2 or logits.shape[1] != self.num_classes:
            raise ValueError(f"expected logits [B,{self.num_classes}], got {tuple(logits.shape)}")
        labels = labels.reshape(-1)
        if labels.dtype != torch.int64:
            labels = labels.long()
        logits = logits.detach().float().contiguous()
        if self.binary:
            ops.metrics_update(logits, labels.contiguous(), self.confusion, self.scores, self.score_labels, self.count)
        else:
            ops.metrics_update(logits, labels.contiguous(), self.confusion)

    def confusion_matrix(self) -> torch.Tensor:
        """int64 [C, C] on the host, rows = target, columns = prediction."""
        c = self.confusion.cpu()
        if int(c[-1]) != 0:
            raise ValueError(f"{int(c[-1])} labels were outside [0, {self.num_classes})")
        return c[:-1].view(self.num_classes, self.num_classes)

    def compute(self) -> Dict[str, float]:
        cm = self.confusion_matrix()
        if not self.binary:
            total = int(cm.sum())
            return {"acc": _safe_divide(int(cm.diag().sum()), total), "confusion": cm}     # micro accuracy (torchmetrics default)
        tn, fp, fn, tp = int(cm[0, 0]), int(cm[0, 1]), int(cm[1, 0]), int(cm[1, 1])
        out = binary_metrics_from_counts(tp, fp, tn, fn)
        n = int(self.count.item())
        if n > self.capacity:
            raise RuntimeError(f"AUROC buffer overflow: {n} samples > capacity {self.capacity}")
        auc = ops.binary_auroc(self.scores, self.score_labels, self.count, self._scratch, self._auc).cpu()
        out["auc"] = float(auc[0])
        out["confusion"] = cm
        return out
