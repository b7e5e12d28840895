"""FusedAdamW: torch.optim.AdamW semantics (lightning_modules.py:599-604, :1108-1113) executed by ONE
multi-tensor libvitk kernel over the model's flat parameter buffer, with Lightning's
clip_grad_norm_(max_norm) (configs/trainer/default.yaml:21,54), the 16-bit weight shadow and the dynamic
loss-scale bookkeeping (skip the step and halve S on fp16 overflow) fused in.

It is a torch.optim.Optimizer (param_groups / state_dict / lr schedulers such as CosineAnnealingLR
keep working: group['lr'] is re-read every step), but step() never touches torch arithmetic.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Union

import torch

from . import ops
from .engine import PAD

CHUNK = 8192  # elements per CTA


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, model, params: Optional[Union[Iterable[torch.nn.Parameter], List[dict]]] = None, lr: float = 1e-3,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2, max_grad_norm: float = 0.0):
        """`model`: a thyroid_vit_cnn_comparison_b200 VisionTransformer/DeiT on a CUDA device.
        `params`: None (all parameters), an iterable of parameters, or torch-style param groups
        (dicts with 'params' and optional 'lr' / 'weight_decay' / 'lr_scale')."""
        self.model = model
        eng = model._ensure_engine()
        self.engine = eng
        import weakref
        eng.fused_optimizer = weakref.ref(self)      # this optimizer owns the fp16 loss-scale bookkeeping from now on
        if params is None:
            params = [p for p in model.parameters()]
        params = list(params)
        if len(params) == 0:
            raise ValueError("optimizer got an empty parameter list")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.max_grad_norm = float(max_grad_norm)
        flat = eng.flat
        dev = flat.params.device
        self.exp_avg = torch.zeros_like(flat.params)
        self.exp_avg_sq = torch.zeros_like(flat.params)
        # state = {step, lr(base, fixed 1.0: per-chunk table carries the real lr), grad_sqnorm, clip_coef}
        self.dev_state = torch.tensor([0.0, 1.0, 0.0, 1.0], dtype=torch.float32, device=dev)
        # partial sums of the fixed-order gradient norm: owned by this optimizer (two optimizers may run on different streams)
        from . import _lib
        self._sq_scratch = torch.zeros(int(_lib.load().vitk_sqnorm_scratch_floats()), dtype=torch.float32, device=dev)
        ptr_to_name = {p.data_ptr(): n for n, p in model._engine_params().items()}
        self._name_of = {}                      # id(parameter) -> flat-buffer entry (parameters the engine owns)
        for g in self.param_groups:
            for p in g["params"]:
                n = ptr_to_name.get(p.data_ptr())
                if n is not None:
                    self._name_of[id(p)] = n
        self._chunk_group: List[int] = []       # group index of every chunk (-1: not optimised -> lr 0)
        offs, lens = [], []
        group_of: Dict[str, int] = {}
        for gi, g in enumerate(self.param_groups):
            b = g.get("betas", betas)
            if tuple(b) != tuple(betas) or g.get("eps", eps) != eps:
                raise NotImplementedError("per-group betas/eps are not supported by the fused kernel")
            for p in g["params"]:
                n = ptr_to_name.get(p.data_ptr())
                if n is not None:
                    group_of[n] = gi        # parameters outside the engine (quality_score: never get a gradient) are ignored
        for n in flat.order:
            off, shape = flat.offsets[n]
            padded = (shape.numel() + PAD - 1) // PAD * PAD
            for c0 in range(0, padded, CHUNK):
                offs.append(off + c0)
                lens.append(min(CHUNK, padded - c0))
                self._chunk_group.append(group_of.get(n, -1))
        self.chunk_off = torch.tensor(offs, dtype=torch.int64, device=dev)
        self.chunk_len = torch.tensor(lens, dtype=torch.int32, device=dev)
        self.chunk_lr = torch.zeros(len(offs), dtype=torch.float32, device=dev)
        self.chunk_wd = torch.zeros(len(offs), dtype=torch.float32, device=dev)
        self._hyper_cache = None
        self.betas, self.eps = tuple(betas), eps
        self._refresh_hyper()

    def _refresh_hyper(self) -> None:
        key = tuple((g["lr"] * g.get("lr_scale_applied", 1.0), g["weight_decay"]) for g in self.param_groups)
        if key == self._hyper_cache:
            return
        lr = [key[gi][0] if gi >= 0 else 0.0 for gi in self._chunk_group]
        wd = [key[gi][1] if gi >= 0 else 0.0 for gi in self._chunk_group]
        self.chunk_lr.copy_(torch.tensor(lr, dtype=torch.float32), non_blocking=False)
        self.chunk_wd.copy_(torch.tensor(wd, dtype=torch.float32), non_blocking=False)
        self._hyper_cache = key

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._refresh_hyper()
        self.launch()
        return loss

    def launch(self) -> None:
        """Enqueue norm + update kernels (CUDA-graph capturable: no host reads, hyper-parameters on device)."""
        flat = self.engine.flat
        # the squared norm doubles as the fp16 overflow detector, so it is always computed
        ops.grad_sqnorm(flat.grads, self.dev_state, self._sq_scratch)
        eng = self.engine
        fp16 = flat.dtype16 == torch.float16
        ops.adamw_step(flat.params, flat.grads, self.exp_avg, self.exp_avg_sq, None if fp16 else flat.w16,
                       flat.w16 if fp16 else None, self.chunk_off, self.chunk_len, self.chunk_lr, self.chunk_wd,
                       self.dev_state, eng.amp if fp16 else None, self.betas[0], self.betas[1], self.eps,
                       self.max_grad_norm, eng.growth_interval)

    def zero_grad(self, set_to_none: bool = True) -> None:
        self.engine.zero_grad()

    # ------------------------------------------------------------------ checkpointing
    # The moments live in two flat buffers and the step count / loss scale on the device, none of which torch's default
    # Optimizer.state_dict() would see.  state_dict() therefore emits exactly torch.optim.AdamW's layout -- per-parameter
    # {'step', 'exp_avg', 'exp_avg_sq'} under the packed parameter index -- so a Lightning checkpoint written by the
    # reference (torch AdamW, lightning_modules.py:599-604) resumes here and vice versa, plus one extra key `vitk` with the
    # device-side scalars torch has no slot for (dynamic loss scale, skipped-step counters).
    def state_dict(self):
        sd = super().state_dict()                    # param_groups with packed indices; `state` is empty (nothing lives there)
        flat = self.engine.flat
        step = self.dev_state[0:1].detach().clone().cpu().reshape(())
        state, idx = {}, 0
        for g in self.param_groups:
            for p in g["params"]:
                n = self._name_of.get(id(p))
                if n is not None and float(step) > 0:
                    state[idx] = {"step": step.clone(), "exp_avg": flat.view(self.exp_avg, n).detach().clone(),
                                  "exp_avg_sq": flat.view(self.exp_avg_sq, n).detach().clone()}
                idx += 1
        sd["state"] = state
        sd["vitk"] = {"amp": self.engine.amp.detach().clone().cpu(), "dev_state": self.dev_state.detach().clone().cpu(),
                      "max_grad_norm": self.max_grad_norm}
        return sd

    def load_state_dict(self, state_dict) -> None:
        sd = dict(state_dict)
        extra = sd.pop("vitk", None)
        super().load_state_dict(sd)                  # validates the group structure, restores lr / weight_decay / ...
        flat = self.engine.flat
        step = None
        for g in self.param_groups:
            for p in g["params"]:
                st = self.state.get(p)
                n = self._name_of.get(id(p))
                if st is None or n is None:
                    continue
                flat.view(self.exp_avg, n).copy_(st["exp_avg"].to(self.exp_avg.device, torch.float32).view_as(p))
                flat.view(self.exp_avg_sq, n).copy_(st["exp_avg_sq"].to(self.exp_avg.device, torch.float32).view_as(p))
                step = float(st["step"]) if step is None else step
        self.state.clear()                           # the flat buffers are the only copy
        if extra is not None:
            self.dev_state.copy_(extra["dev_state"].to(self.dev_state.device))
            self.engine.amp.copy_(extra["amp"].to(self.engine.amp.device))
        elif step is not None:                       # a plain torch.optim.AdamW checkpoint: only the step count carries over
            self.dev_state[0] = step
        self._hyper_cache = None
        self._refresh_hyper()

    @property
    def last_clip_coef(self) -> float:
        return float(self.dev_state[3].item())
