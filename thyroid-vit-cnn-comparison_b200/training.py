"""Training-step layer: the B200 counterparts of the reference's LightningModules
(src/training/lightning_modules.py) and a graph-capturable native step.

  fused_cross_entropy / fused_distillation_loss   autograd-aware wrappers of the fused loss kernel
  ThyroidViTModule                                 lightning_modules.py:310-731   (training_step :441-473,
                                                   configure_optimizers :576-626, _get_parameter_groups_with_decay :628-659)
  ThyroidDistillationModule                        lightning_modules.py:742-1160  (training_step :949-988,
                                                   get_current_alpha :922-938, configure_optimizers :1084-1147)
  TrainStep                                        the whole step (H2D -> fwd -> loss -> bwd -> [all-reduce] -> clip ->
                                                   AdamW) as one enqueue, optionally replayed from a CUDA graph

Metric names logged are the reference's: train_loss, train_acc, class_loss, distill_loss, alpha,
teacher_agreement, val_loss, val_acc.
"""
from __future__ import annotations

import json
import math
from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from . import teacher as _teacher
from .optim import FusedAdamW

try:
    import pytorch_lightning as _pl
    _Base = _pl.LightningModule
except Exception:  # pragma: no cover
    _pl = None
    _Base = nn.Module


# --------------------------------------------------------------------------- fused losses (autograd)
class _FusedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cls_logits, dist_logits, teacher_logits, labels, mode, w_cls, w_dist, T, ls, grad_div):
        out, dcls, ddist = ops.loss_fwd_bwd(cls_logits.detach().contiguous().float(),
                                            None if dist_logits is None else dist_logits.detach().contiguous().float(),
                                            None if teacher_logits is None else teacher_logits.detach().contiguous().float(),
                                            labels, mode=mode, w_cls=w_cls, w_dist=w_dist, T=T, label_smoothing=ls,
                                            grad_div=grad_div)
        ctx.save_for_backward(dcls, ddist if ddist is not None else dcls)
        ctx.has_dist = dist_logits is not None
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, gloss, _gstats):
        dcls, ddist = ctx.saved_tensors
        g0 = dcls * gloss
        g1 = ddist * gloss if ctx.has_dist else None
        return g0, g1, None, None, None, None, None, None, None, None


def _labels(labels: torch.Tensor) -> torch.Tensor:
    """lightning_modules.py:445-451: long dtype, squeeze a trailing singleton."""
    if labels.dtype != torch.long:
        labels = labels.long()
    if labels.dim() == 0:
        labels = labels.unsqueeze(0)
    elif labels.dim() > 1 and labels.shape[-1] == 1:
        labels = labels.squeeze(-1)
    return labels.contiguous()


def fused_cross_entropy(outputs, labels: torch.Tensor, label_smoothing: float = 0.0, grad_div: float = 1.0):
    """CE, or 0.5*CE(cls)+0.5*CE(dist) for a (cls, dist) tuple -- lightning_modules.py:453-465.
    Returns (loss, {'acc', 'n_correct'}) with the counters still on the device."""
    labels = _labels(labels)
    if isinstance(outputs, tuple) and len(outputs) == 2:
        loss, st = _FusedLossFn.apply(outputs[0], outputs[1], None, labels, 0, 0.5, 0.5, 1.0, label_smoothing, grad_div)
    else:
        loss, st = _FusedLossFn.apply(outputs, None, None, labels, 0, 1.0, 0.0, 1.0, label_smoothing, grad_div)
    return loss, {"acc": st[3] / st[5], "n_correct": st[3]}


def fused_distillation_loss(student_outputs, labels: torch.Tensor, teacher_logits: torch.Tensor, alpha: float,
                            temperature: float, distillation_type: str = "soft", label_smoothing: float = 0.0,
                            grad_div: float = 1.0):
    """(1-alpha)*CE(cls,y) + alpha*[KL_T(dist||teacher)*T^2 | CE(dist, argmax teacher)] -- lightning_modules.py:949-974.
    Returns (total, stats) with stats = {class_loss, distill_loss, acc, teacher_agreement}."""
    labels = _labels(labels)
    if isinstance(student_outputs, tuple) and len(student_outputs) == 2:
        c, d = student_outputs
    else:
        c = d = student_outputs                              # :957 single-logit student: cls = dist
    mode = 1 if distillation_type == "soft" else 2
    if d is c:
        d = c.view_as(c)                                     # distinct autograd edge, same logits: gradients add
    total, st = _FusedLossFn.apply(c, d, teacher_logits, labels, mode, 1.0 - alpha, alpha, temperature, label_smoothing, grad_div)
    return total, {"class_loss": st[1], "distill_loss": st[2], "acc": st[3] / st[5], "teacher_agreement": st[4] / st[5]}


class DistillationLoss(nn.Module):
    """deit_models.py:417-480, same constructor; forward runs the fused kernel."""

    def __init__(self, base_criterion: nn.Module = None, teacher_model: Optional[nn.Module] = None,
                 distillation_type: str = "soft", alpha: float = 0.5, tau: float = 3.0):
        super().__init__()
        self.base_criterion = base_criterion if base_criterion is not None else nn.CrossEntropyLoss()
        self.teacher_model = teacher_model
        self.distillation_type = distillation_type
        self.alpha = alpha
        self.tau = tau

    def forward(self, outputs, targets, teacher_outputs=None):
        ls = float(getattr(self.base_criterion, "label_smoothing", 0.0))
        if isinstance(outputs, tuple):
            outputs_cls, outputs_dist = outputs
        else:
            outputs_cls, outputs_dist = outputs, None
        if outputs_dist is None or teacher_outputs is None:
            return fused_cross_entropy(outputs_cls, targets, ls)[0]               # :461-465
        return fused_distillation_loss((outputs_cls, outputs_dist), targets, teacher_outputs, self.alpha, self.tau,
                                       self.distillation_type, ls)[0]


# --------------------------------------------------------------------------- config helpers
def cfg_get(cfg: Any, key: str, default=None):
    """Attribute-or-.get access over OmegaConf / dict / plain objects (tests/unit/test_models.py:10-22 contract)."""
    if cfg is None:
        return default
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    if hasattr(cfg, key):
        v = getattr(cfg, key)
        return default if v is None else v
    g = getattr(cfg, "get", None)
    if callable(g):
        try:
            return g(key, default)
        except Exception:
            return default
    return default


def _split_metrics(module, split: str):
    from .metrics import ClassificationMetrics
    store = module.__dict__.setdefault("_metrics", {})
    if split not in store:
        num_classes = int(cfg_get(cfg_get(module.config, "dataset"), "num_classes", 2) or 2)
        store[split] = ClassificationMetrics(num_classes)
    return store[split]


def _eval_step(module, model, batch, split: str):
    """validation_step / test_step of both Lightning modules: loss + the metric updates that torchmetrics does on the
    host in the reference, here one libvitk launch on device counters (metrics.ClassificationMetrics).  The running
    values the reference logs every step (`self.log('val_acc', self.val_acc, ...)`) are exposed through
    `split_metrics(split).compute()` at epoch end instead, so the step itself never synchronises with the host."""
    images, labels = batch
    outputs = model(images)
    if isinstance(outputs, tuple):
        outputs = outputs[0]
    loss, stats = fused_cross_entropy(outputs, labels, getattr(module, "label_smoothing", 0.0))
    _split_metrics(module, split).update(outputs, _labels(labels))
    module.log(f"{split}_loss", loss, on_step=False, on_epoch=True, prog_bar=True)
    module.log(f"{split}_acc", stats["acc"], on_step=False, on_epoch=True, prog_bar=True)
    return {f"{split}_loss": loss, f"{split}_acc": stats["acc"], "preds": outputs.argmax(dim=1)}


class ThyroidViTModule(_Base):
    """Counterpart of lightning_modules.py:310-731 for the B200 path."""

    def __init__(self, config, optimizer_params: Optional[Dict] = None, model: Optional[nn.Module] = None):
        super().__init__()
        self.config = config
        if optimizer_params is None:                                              # :329-338
            try:
                with open("configs/vit_optimizer_params.json", "r") as f:
                    optimizer_params = json.load(f)
            except Exception:
                optimizer_params = cfg_get(cfg_get(config, "training"), "optimizer_params")
                if optimizer_params is None:
                    raise ValueError("Optimizer parameters not provided and failed to load from JSON")
        self.optimizer_params = dict(optimizer_params)
        self.model = model if model is not None else self._create_model()
        loss_cfg = cfg_get(config, "loss")
        self.label_smoothing = float(cfg_get(loss_cfg, "label_smoothing", 0.0))   # :343-350
        self.num_classes = cfg_get(cfg_get(config, "dataset"), "num_classes", 2)
        self.max_grad_norm = float(cfg_get(cfg_get(config, "trainer"), "gradient_clip_val", 0.0) or 0.0)
        self.logged: Dict[str, Any] = {}
        self._metrics: Dict[str, Any] = {}

    def _create_model(self) -> nn.Module:                                         # :379-400
        from .registry import ModelRegistry
        model_cfg = cfg_get(self.config, "model")
        if model_cfg is None:
            raise AttributeError("ThyroidViTModule requires 'config.model' for ModelRegistry.")
        return ModelRegistry.create_model(model_cfg)

    if _pl is None:
        def log(self, name, value, **kw):
            self.logged[name] = value

    def forward(self, x):
        return self.model(x)

    def training_step(self, batch, batch_idx):                                    # :441-473
        images, labels = batch
        outputs = self.model(images)
        loss, stats = fused_cross_entropy(outputs, labels, self.label_smoothing)
        self.log("train_loss", loss, on_step=True, on_epoch=True, prog_bar=True)
        self.log("train_acc", stats["acc"], on_step=True, on_epoch=True, prog_bar=True)
        return loss

    def validation_step(self, batch, batch_idx):                                  # :474-520
        return _eval_step(self, self.model, batch, "val")

    def test_step(self, batch, batch_idx):                                        # :522-570
        return _eval_step(self, self.model, batch, "test")

    def split_metrics(self, split: str = "val"):
        """The on-device accumulator of one split (metrics.ClassificationMetrics); compute() / reset() at epoch end."""
        return _split_metrics(self, split)

    def _inner(self):
        m = self.model
        return getattr(m, "model", m) if not hasattr(m, "_ensure_engine") else m

    def configure_optimizers(self):                                               # :576-626
        op = self.optimizer_params
        base_lr = float(op.get("lr", 0.001))
        weight_decay = float(op.get("weight_decay", 0.05))
        betas = tuple(op.get("betas", (0.9, 0.999)))
        inner = self._inner()
        tr = cfg_get(self.config, "training")
        if cfg_get(tr, "layer_wise_lr_decay", False):
            groups = self._get_parameter_groups_with_decay(base_lr)
        else:
            groups = None
        opt = FusedAdamW(inner, groups, lr=base_lr, betas=betas, weight_decay=weight_decay, max_grad_norm=self.max_grad_norm)
        sched = cfg_get(tr, "scheduler_params")
        if sched is not None:
            name = str(cfg_get(sched, "name", "cosineannealinglr")).lower()
            if name == "cosineannealinglr":
                scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(
                    opt, T_max=int(cfg_get(sched, "T_max", cfg_get(tr, "epochs", 100))), eta_min=float(cfg_get(sched, "eta_min", 1e-6)))
                return {"optimizer": opt, "lr_scheduler": {"scheduler": scheduler, "interval": "epoch"}}
            return opt
        return opt

    def _get_parameter_groups_with_decay(self, base_lr: float):                   # :628-659
        tr = cfg_get(self.config, "training")
        decay = float(cfg_get(cfg_get(tr, "layer_decay", {}), "decay_rate", 0.75))
        m = self._inner()
        groups = []
        embed = list(m.patch_embed.parameters()) + [m.cls_token, m.pos_embed]
        groups.append({"params": embed, "lr": base_lr * decay ** 2})
        n = len(m.blocks)
        for i, blk in enumerate(m.blocks):
            groups.append({"params": list(blk.parameters()), "lr": base_lr * decay ** (n - i - 1)})
        groups.append({"params": list(m.head.parameters()), "lr": base_lr})
        return groups


class ThyroidDistillationModule(_Base):
    """Counterpart of lightning_modules.py:742-1160: frozen teacher forward + student step + fused KL/CE."""

    def __init__(self, config, trainer=None, student: Optional[nn.Module] = None, teacher: Optional[nn.Module] = None):
        super().__init__()
        self.config = config
        dist = cfg_get(config, "distillation")
        if not cfg_get(dist, "enabled", False):
            raise ValueError("Distillation must be enabled in config")                      # :761-762
        self.student = student if student is not None else self._create_student_model()
        if teacher is None:
            raise ValueError("a teacher nn.Module must be supplied (checkpoint loading is outside the accelerated path)")
        self.teacher = teacher
        if cfg_get(dist, "freeze_teacher", True):                                           # :771-773
            for p in self.teacher.parameters():
                p.requires_grad = False
        self.alpha = float(cfg_get(dist, "alpha", 0.5))
        self.temperature = float(cfg_get(dist, "temperature", 4.0))
        self.distillation_type = cfg_get(dist, "distillation_type", "soft")
        loss_cfg = cfg_get(config, "loss") or cfg_get(cfg_get(config, "training"), "loss")
        self.label_smoothing = float(cfg_get(loss_cfg, "label_smoothing", 0.0))
        self.criterion = DistillationLoss(nn.CrossEntropyLoss(label_smoothing=self.label_smoothing), None,
                                          self.distillation_type, self.alpha, self.temperature)
        self.progressive_schedule = None
        if cfg_get(dist, "progressive_distillation", False):
            sch = cfg_get(dist, "progressive_schedule", {}) or {}
            self.progressive_schedule = {int(k): float(v) for k, v in dict(sch).items()}
        self.max_grad_norm = float(cfg_get(cfg_get(config, "trainer"), "gradient_clip_val", 0.0) or 0.0)
        self._epoch = 0
        self.logged: Dict[str, Any] = {}

    def _create_student_model(self) -> nn.Module:                                          # :807-832
        from .registry import ModelRegistry
        cfg = cfg_get(self.config, "student_model") or cfg_get(self.config, "model")
        if cfg is None:
            raise AttributeError("Student model configuration not found in cfg.student_model or cfg.model")
        return ModelRegistry.create_model(cfg)

    if _pl is None:
        def log(self, name, value, **kw):
            self.logged[name] = value

        @property
        def current_epoch(self) -> int:
            return self._epoch

    def get_current_alpha(self) -> float:                                                  # :922-938
        if not self.progressive_schedule:
            return self.alpha
        a = self.alpha
        for thr in sorted(self.progressive_schedule):
            if self.current_epoch >= thr:
                a = self.progressive_schedule[thr]
            else:
                break
        return a

    def forward(self, x):
        return self.student(x)

    def get_teacher_outputs(self, x: torch.Tensor) -> torch.Tensor:                         # :943-947
        self.teacher.eval()
        with torch.no_grad():
            return self.teacher(x)

    def training_step(self, batch, batch_idx):                                              # :949-988
        images, labels = batch
        teacher_outputs = self.get_teacher_outputs(images).float()
        student_outputs = self.student(images)
        alpha = self.get_current_alpha()
        total, st = fused_distillation_loss(student_outputs, labels, teacher_outputs, alpha, self.temperature,
                                            self.distillation_type, self.label_smoothing)
        self.log("train_loss", total, on_step=True, on_epoch=True, prog_bar=True)
        self.log("train_acc", st["acc"], on_step=True, on_epoch=True, prog_bar=True)
        self.log("class_loss", st["class_loss"], on_step=False, on_epoch=True)
        self.log("distill_loss", st["distill_loss"], on_step=False, on_epoch=True)
        self.log("alpha", alpha, on_step=False, on_epoch=True)
        self.log("teacher_agreement", st["teacher_agreement"], on_step=False, on_epoch=True)
        return total

    def validation_step(self, batch, batch_idx):                                            # :990-1040
        out = _eval_step(self, self.student, batch, "val")
        images, _ = batch
        teacher_preds = self.get_teacher_outputs(images).argmax(dim=1)
        agreement = (out["preds"] == teacher_preds).float().mean()                          # :1022-1024
        self.log("val_teacher_agreement", agreement, on_step=False, on_epoch=True)
        out["val_teacher_agreement"] = agreement
        return out

    def test_step(self, batch, batch_idx):                                                  # :1042-1082
        return _eval_step(self, self.student, batch, "test")

    def split_metrics(self, split: str = "val"):
        return _split_metrics(self, split)

    def configure_optimizers(self):                                                         # :1084-1147
        tr = cfg_get(self.config, "training")
        op = cfg_get(tr, "optimizer_params")
        if op is None:
            raise ValueError("Missing optimizer_params in training configuration for ThyroidDistillationModule.")
        base_lr = float(cfg_get(op, "lr", 0.001))
        weight_decay = float(cfg_get(op, "weight_decay", 0.05))
        betas = tuple(cfg_get(op, "betas", (0.9, 0.999)))
        student = self.student if hasattr(self.student, "_ensure_engine") else getattr(self.student, "model")
        groups = student.get_parameter_groups(weight_decay=weight_decay)                    # :1095-1099
        for g in groups:
            if "lr" not in g:
                g["lr"] = base_lr * g.get("lr_scale", 1.0)                                  # :1101-1103
        opt = FusedAdamW(student, groups, lr=base_lr, betas=betas, weight_decay=weight_decay, max_grad_norm=self.max_grad_norm)
        sched = cfg_get(tr, "scheduler_params")
        if sched is not None and str(cfg_get(sched, "name", "cosineannealinglr")).lower() == "cosineannealinglr":
            scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(
                opt, T_max=int(cfg_get(sched, "T_max", cfg_get(tr, "epochs", 100))), eta_min=float(cfg_get(sched, "eta_min", 1e-6)))
            return {"optimizer": opt, "lr_scheduler": {"scheduler": scheduler, "interval": "epoch", "frequency": 1}}
        return opt


# --------------------------------------------------------------------------- native whole-step
class TrainStep:
    """One full training step enqueued without autograd or per-op Python state:

        [H2D images, labels] -> zero flat grads -> encoder fwd -> fused loss (+grad) -> encoder bwd
        -> [bucketed all-reduce] -> grad-norm + clip + AdamW + bf16 shadow

    `mode`: 'ce' (ThyroidViTModule.training_step) or 'distill' (ThyroidDistillationModule.training_step; the frozen
    teacher runs under no_grad before the student).  With `use_graph=True` the device work is captured once into a
    CUDA graph (static input buffers) and replayed.  Gradients are pre-divided by `world_size`, so a SUM all-reduce
    yields the global-batch mean exactly as DDP would.
    """

    def __init__(self, model, optimizer: FusedAdamW, batch_size: int, *, mode: str = "ce", teacher: Optional[nn.Module] = None,
                 alpha: float = 0.7, temperature: float = 3.0, distillation_type: str = "soft", label_smoothing: float = 0.0,
                 reducer=None, use_graph: bool = False, teacher_dtype=torch.float32, teacher_fast: bool = True,
                 input_format: str = "nchw", tile_dtype=torch.uint16, tile_mean=None, tile_std=None):
        """teacher_dtype: torch.float32 (default) calls the frozen teacher exactly as the reference does
        (lightning_modules.py:943-947, fp32 eval forward).  torch.bfloat16 / float16 is the OPT-IN fast path: the images are
        cast to that type, channels_last, and a torchvision-style DenseNet runs through teacher.FrozenDenseNet with
        BatchNorm-folded 16-bit weights; its logits differ from the fp32 teacher's by the 16-bit rounding (about 1e-2
        relative, bounded in tests/test_model_gpu.py::test_frozen_densenet_teacher_fast_path_matches_module).
        input_format: 'nchw' -- batches are float [B, chans, H, W] as the reference's DataLoader emits them;
        'gray' -- batches are the single-channel tiles BEFORE the loader's `x.repeat(3,1,1)` + T.Normalize
        (vit_transforms.py:381-393), [B,H,W] / [B,1,H,W] of `tile_dtype` (raw uint16 as stored in the CARS TIFFs, /65535 on
        the device -- dataset.py:549 --, or fp16 / fp32 in [0,1]); the channel replication and the optional
        Normalize(tile_mean, tile_std) are fused into the kernel that writes the patch matrix, so a step uploads
        2 bytes per pixel instead of 12."""
        self.model, self.opt, self.B, self.mode = model, optimizer, batch_size, mode
        self.teacher, self.alpha, self.T = teacher, alpha, temperature
        self.distillation_type, self.ls = distillation_type, label_smoothing
        self.reducer = reducer
        self.world = reducer.world_size if reducer is not None else 1
        self.eng = model._ensure_engine()
        d = self.eng.d
        dev = self.eng.device
        # two input slots: the pinned-host -> HBM copy of step i+1 runs on its own stream (copy engine) while step i computes
        if input_format not in ("nchw", "gray"):
            raise ValueError("input_format must be 'nchw' or 'gray'")
        self.input_format = input_format
        self._gray = None
        if input_format == "gray":
            from .engine import GraySpec
            self._gray = GraySpec(mean=None if tile_mean is None else tuple(tile_mean), std=None if tile_std is None else tuple(tile_std))
            self._images = [torch.zeros(batch_size, d.img, d.img, dtype=tile_dtype, device=dev) for _ in range(2)]
        else:
            self._images = [torch.zeros(batch_size, d.chans, d.img, d.img, dtype=torch.float32, device=dev) for _ in range(2)]
        self._labels = [torch.zeros(batch_size, dtype=torch.int64, device=dev) for _ in range(2)]
        self.slot = 0
        self.stats = torch.zeros(8, dtype=torch.float32, device=dev)
        self.teacher_dtype = teacher_dtype
        self.graphs = [None, None]
        self.use_graph = use_graph
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._teacher_stream = torch.cuda.Stream(device=dev)
        self._copied = [torch.cuda.Event() for _ in range(2)]     # slot filled (recorded on the copy stream)
        self._consumed = [torch.cuda.Event() for _ in range(2)]   # slot no longer read (recorded on the compute stream)
        self._pending_copy = False
        if mode == "distill" and teacher is None:
            raise ValueError("mode='distill' needs a teacher module")
        # a frozen torchvision-style DenseNet teacher runs through the fused eval-mode executor (teacher.py): no torch.cat, one
        # pass for BatchNorm + ReLU, BatchNorm folded into the convolutions; any other teacher module is called as it is
        self._teacher_fn = teacher
        if (teacher is not None and teacher_fast and teacher_dtype in (torch.bfloat16, torch.float16)
                and _teacher.is_supported(teacher)):
            teacher.eval()                                                                   # lightning_modules.py:944
            self._teacher_fn = _teacher.FrozenDenseNet(teacher, dtype=teacher_dtype).to(dev)
        if reducer is not None:
            reducer.attach(self.eng)

    @property
    def images(self) -> torch.Tensor:
        return self._images[self.slot]

    @property
    def labels(self) -> torch.Tensor:
        return self._labels[self.slot]

    def _device_step(self) -> None:
        eng = self.eng
        model = self.model
        teacher_logits = None
        if self.mode == "distill":
            # the frozen teacher (cuDNN, reference: lightning_modules.py:943-947) only feeds the loss: it runs on a side stream,
            # concurrently with the student's forward (a parallel branch of the captured graph), and is joined before the loss
            cur = torch.cuda.current_stream()
            side = self._teacher_stream
            side.wait_stream(cur)
            with torch.cuda.stream(side), torch.no_grad():
                x = self.images
                if self._gray is not None:       # the teacher module wants the loader's [B,C,H,W] batch: built on the side stream
                    g = ops.resize_u16(x, eng.d.img, eng.d.img) if x.dtype == torch.uint16 else x.float()
                    x = ops.finish_tiles(g, eng.d.chans, mean=self._gray.mean, std=self._gray.std)
                if self.teacher_dtype is not None and self.teacher_dtype != torch.float32:
                    x = x.to(self.teacher_dtype).contiguous(memory_format=torch.channels_last)
                teacher_logits = self._teacher_fn(x).float().contiguous()
        eng.zero_grad()
        l0, l1 = eng.forward(self.images, train=True, gray=self._gray)
        eng.generation += 1
        if self.mode == "distill":
            torch.cuda.current_stream().wait_stream(self._teacher_stream)
        if self.mode == "ce":
            if l1 is not None:
                out, d0, d1 = ops.loss_fwd_bwd(l0, l1, None, self.labels, mode=0, w_cls=0.5, w_dist=0.5,
                                               label_smoothing=self.ls, grad_div=float(self.world))
            else:
                out, d0, d1 = ops.loss_fwd_bwd(l0, None, None, self.labels, mode=0, w_cls=1.0, w_dist=0.0,
                                               label_smoothing=self.ls, grad_div=float(self.world))
        else:
            md = 1 if self.distillation_type == "soft" else 2
            if l1 is None:
                raise NotImplementedError("TrainStep distillation expects a distilled student (cls + dist logits)")
            out, d0, d1 = ops.loss_fwd_bwd(l0, l1, teacher_logits, self.labels, mode=md, w_cls=1.0 - self.alpha, w_dist=self.alpha,
                                           T=self.T, label_smoothing=self.ls, grad_div=float(self.world))
        self.stats.copy_(out)
        eng.backward(self.B, d0, d1)
        if self.reducer is not None:
            self.reducer.finish()
        self.opt.launch()

    def load(self, images: torch.Tensor, labels: torch.Tensor) -> None:
        """Host (ideally pinned) or device tensors -> the next input slot.  Host batches are copied on a dedicated
        stream: the copy starts as soon as the slot's previous reader (two steps back) is done, i.e. it overlaps the
        step that is still computing; run() makes the compute stream wait for it."""
        self.slot ^= 1
        s = self.slot
        cur = torch.cuda.current_stream()
        if self._gray is not None and images.dim() == 4 and images.shape[1] == 1:
            images = images[:, 0]
        if images.is_cuda:
            self._images[s].copy_(images, non_blocking=True)
            self._labels[s].copy_(labels, non_blocking=True)
            self._pending_copy = False
            return
        cs = self._copy_stream
        cs.wait_event(self._consumed[s])
        with torch.cuda.stream(cs):
            self._images[s].copy_(images, non_blocking=True)
            self._labels[s].copy_(labels, non_blocking=True)
            self._copied[s].record(cs)
        self._pending_copy = True

    def run(self) -> None:
        """Enqueue one step on the current stream (no host sync)."""
        model = self.model
        # parameters edited through torch since the last step (load_state_dict, manual copy_) invalidate the 16-bit shadow the
        # kernels read; the fused optimizer itself keeps it in sync without bumping tensor versions, so this is a no-op in
        # the steady state
        model._sync_shadow()
        self.opt._refresh_hyper()
        cur = torch.cuda.current_stream()
        if self._pending_copy:
            cur.wait_event(self._copied[self.slot])
            self._pending_copy = False
        if not self.use_graph:
            self._device_step()
            self._consumed[self.slot].record(cur)
            return
        if self.graphs[self.slot] is None:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):           # warm-up outside capture (lazy inits, workspace allocation)
                snapshot = self._snapshot()
                self._device_step()
                self._restore(snapshot)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()            # one graph per input slot (the slot's buffers are baked into it)
            with torch.cuda.graph(g):
                self._device_step()
            self.graphs[self.slot] = g
            self._restore(snapshot)               # capture does not execute, but keep state exact anyway
        self.graphs[self.slot].replay()
        self._consumed[self.slot].record(cur)

    def _snapshot(self):
        f, o = self.eng.flat, self.opt
        return (f.params.clone(), f.w16.clone(), o.exp_avg.clone(), o.exp_avg_sq.clone(), o.dev_state.clone(), self.eng.amp.clone())

    def _restore(self, snap) -> None:
        f, o = self.eng.flat, self.opt
        f.params.copy_(snap[0]); f.w16.copy_(snap[1]); o.exp_avg.copy_(snap[2]); o.exp_avg_sq.copy_(snap[3])
        o.dev_state.copy_(snap[4]); self.eng.amp.copy_(snap[5])

    def __call__(self, images: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """Public step: copies the batch in, runs the step, returns the device tensor of step statistics
        [loss, class_loss, dist_loss, n_correct, n_agree, B, 0, 0] (read it with .cpu() when needed)."""
        self.load(images, labels)
        self.run()
        return self.stats
