"""Drop-in model classes: same names, constructor kwargs, attributes and state_dict keys as the
reference's hand-written ViT/DeiT (src/models/vit/vision_transformer_base.py, vit_models.py,
deit_models.py), but every forward/backward runs in libvitk.so through `engine.VitEngine`.

The nn.Module tree below only OWNS parameters (so state_dict()/load_state_dict()/optimizers/
Lightning see exactly the reference's layout); it contains no torch arithmetic.  There is no CPU
path: calling a model whose parameters are not on a CUDA device raises.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .engine import Dims, VitEngine

try:  # if Lightning is installed the classes plug into pl.Trainer like the reference's do
    import pytorch_lightning as _pl
    _Base = _pl.LightningModule
except Exception:  # pragma: no cover - Lightning is absent in the build image
    _pl = None
    _Base = nn.Module


class _HParams(dict):
    """hparams container supporting both attribute access and .get (vision_transformer_base.py:433,564)."""
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


def get_layer_from_string(layer_name):
    """vision_transformer_base.py:20-46: string layer names -> classes."""
    if layer_name is None:
        return None
    if not isinstance(layer_name, str):
        return layer_name
    layer_map = {
        "LayerNorm": nn.LayerNorm, "nn.LayerNorm": nn.LayerNorm, "GELU": nn.GELU, "nn.GELU": nn.GELU,
        "ReLU": nn.ReLU, "nn.ReLU": nn.ReLU, "SiLU": nn.SiLU, "nn.SiLU": nn.SiLU,
        "Identity": nn.Identity, "nn.Identity": nn.Identity,
    }
    if layer_name in layer_map:
        return layer_map[layer_name]
    raise ValueError(f"Unknown layer name: {layer_name}")


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: this model runs only on a CUDA device through libvitk.so "
                           "(sm_100a kernels); there is no CPU fallback -- call .cuda() first")


class DropPath(nn.Module):
    """Stochastic depth container (vision_transformer_base.py:49-64).  Identity in eval / rate 0."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        # inside a model the engine applies stochastic depth in the residual epilogue of the branch's last GEMM
        # (VitEngine.set_drop_path); called stand-alone on a tensor only the identity cases are served
        if self.drop_prob == 0.0 or not self.training:
            return x
        # stand-alone call (outside a model): the per-sample factors floor(keep + u) / keep come from the same kernel the
        # engine uses (vitk_droppath_scale); only the broadcast multiply -- differentiable, off the hot path -- is torch's
        _require_cuda(x, "DropPath")
        u = torch.rand(1, x.shape[0], dtype=torch.float32, device=x.device)
        prob = torch.tensor([float(self.drop_prob)], dtype=torch.float32, device=x.device)
        scale = ops.droppath_scale(u, prob, 1)[0]
        return x * scale.view((x.shape[0],) + (1,) * (x.ndim - 1)).to(x.dtype)


class PatchEmbed(nn.Module):
    """Parameter container for the patch projection (vision_transformer_base.py:67-143).
    `proj` keeps the Conv2d weight layout [D, C, P, P]; `quality_score` keeps the reference's (dead)
    quality branch parameters so checkpoints round-trip."""

    def __init__(self, img_size: int = 256, patch_size: int = 16, in_chans: int = 1, embed_dim: int = 768,
                 norm_layer=None, flatten: bool = True, bias: bool = True, strict_img_size: bool = True,
                 projection_type: str = "conv", quality_aware: bool = True):
        super().__init__()
        self.img_size, self.patch_size = img_size, patch_size
        self.grid_size = img_size // patch_size
        self.num_patches = self.grid_size ** 2
        self.flatten, self.projection_type = flatten, projection_type
        self.strict_img_size, self.quality_aware = strict_img_size, quality_aware
        self.in_chans, self.embed_dim = in_chans, embed_dim
        if projection_type == "conv":
            self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)
        else:   # 'linear' (:102-107): channel-last patch vectors + nn.Linear; index 0 holds no parameters, as in the reference
            from einops.layers.torch import Rearrange
            self.proj = nn.Sequential(Rearrange("b c (h p1) (w p2) -> b (h w) (p1 p2 c)", p1=patch_size, p2=patch_size),
                                      nn.Linear(patch_size * patch_size * in_chans, embed_dim, bias=bias))
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()
        if quality_aware:
            self.quality_score = nn.Sequential(nn.Conv2d(in_chans, 32, kernel_size=3, padding=1), nn.ReLU(inplace=True),
                                               nn.Conv2d(32, 1, kernel_size=1), nn.Sigmoid())

    def forward(self, x: torch.Tensor):
        """Standalone (inference-only) patch projection -> ([B, num_patches, D] fp32, quality scores [B, num_patches] | None).
        The quality branch (vision_transformer_base.py:126-132) is dead code inside the model's forward -- its scores never
        reach the tokens -- so the engine skips it; a stand-alone call evaluates it with torch's convolutions (two tiny
        convs + pooling, not on the training path) so callers of the reference's PatchEmbed API get the same tuple."""
        B, C, H, W = x.shape
        if self.strict_img_size:
            assert H == self.img_size and W == self.img_size, \
                f"Input size ({H}x{W}) doesn't match expected size ({self.img_size}x{self.img_size})"
        _require_cuda(x, "PatchEmbed")
        linear = self.projection_type != "conv"
        lin = self.proj[1] if linear else self.proj
        patches = ops.patchify(x.float().contiguous(), self.patch_size, channel_last=linear)
        w16 = ops.cast_fp16(lin.weight.detach().reshape(self.embed_dim, -1).contiguous())
        out = torch.empty(patches.shape[0], self.embed_dim, dtype=torch.float32, device=x.device)
        ops.gemm(patches, w16, patches.shape[0], self.embed_dim, patches.shape[1], out=out,
                 bias=lin.bias.detach() if lin.bias is not None else None)
        quality = None
        if self.quality_aware and hasattr(self, "quality_score"):
            with torch.no_grad():
                quality = torch.nn.functional.avg_pool2d(self.quality_score(x.float()), self.patch_size).flatten(1)
        return out.view(B, -1, self.embed_dim), quality


class Attention(nn.Module):
    """Parameter container for qkv/proj (vision_transformer_base.py:146-195)."""

    def __init__(self, dim: int, num_heads: int = 8, qkv_bias: bool = False, attn_drop: float = 0.0,
                 proj_drop: float = 0.0, store_attention: bool = True):
        super().__init__()
        assert dim % num_heads == 0, f"dim {dim} must be divisible by num_heads {num_heads}"
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.store_attention = store_attention
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self._attention_maps = None

    @property
    def attention_maps(self):
        """CPU copy of the last eval-mode attention probabilities [B,H,N,N] (reference :186-188),
        materialised lazily so the D2H copy never sits on the forward path."""
        m = self._attention_maps
        if m is not None and m.is_cuda:
            m = m.detach().cpu()
            self._attention_maps = m
        return m

    @attention_maps.setter
    def attention_maps(self, v):
        self._attention_maps = v


class Mlp(nn.Module):
    """Parameter container for fc1/fc2 (vision_transformer_base.py:198-223); activation is exact-erf GELU."""

    def __init__(self, in_features: int, hidden_features: Optional[int] = None, out_features: Optional[int] = None,
                 act_layer=nn.GELU, drop: float = 0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        act_layer = get_layer_from_string(act_layer) or nn.GELU
        if act_layer is not nn.GELU:
            raise NotImplementedError("the fused fc1 epilogue implements exact-erf GELU only (reference default)")
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)


class Block(nn.Module):
    """Pre-norm transformer block container (vision_transformer_base.py:226-285)."""

    def __init__(self, dim: int, num_heads: int, mlp_ratio: float = 4.0, qkv_bias: bool = False, drop: float = 0.0,
                 attn_drop: float = 0.0, drop_path: float = 0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 store_attention: bool = True):
        super().__init__()
        norm_layer = get_layer_from_string(norm_layer) or nn.LayerNorm
        if getattr(norm_layer, "func", norm_layer) is not nn.LayerNorm:     # functools.partial(nn.LayerNorm, eps=...) is fine
            raise NotImplementedError("only nn.LayerNorm is implemented in the sm_100a path")
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop,
                              store_attention=store_attention)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)


class _EncoderFn(torch.autograd.Function):
    """Whole-encoder autograd node.  Parameter gradients are accumulated by libvitk directly into the
    engine's flat gradient buffer (the tensors `param.grad` are views of it)."""

    @staticmethod
    def forward(ctx, model, images, anchor):
        eng = model._engine
        l0, l1 = eng.forward(images, train=True)
        eng.generation += 1
        ctx.model, ctx.B, ctx.gen = model, images.shape[0], eng.generation
        ctx.two = l1 is not None
        if l1 is None:
            return l0
        return l0, l1

    @staticmethod
    def backward(ctx, *grads):
        model = ctx.model
        eng = model._engine
        if ctx.gen != eng.generation:
            raise RuntimeError("backward through a stale forward: the activation workspace holds a later forward "
                               "(run forward/backward pairs in order)")
        model._bind_grads()
        dl0 = grads[0].contiguous().float()
        dl1 = grads[1].contiguous().float() if ctx.two else None
        eng.backward(ctx.B, dl0, dl1)
        if eng.fused_optimizer is None or eng.fused_optimizer() is None:
            # an EXTERNAL torch optimizer will read param.grad: overflow check (zeroes the gradients of an overflowed step)
            # + loss-scale bookkeeping happen here.  With a FusedAdamW attached, its own norm pass is the overflow detector
            # and its tick kernel the only owner of the loss-scale state (one bookkeeping update per optimizer step).
            eng.amp_update()
        return None, None, None


class _FeaturesFn(torch.autograd.Function):
    """Autograd node of a training-mode forward_features(): encoder + final norm + pooling + pre_logits, no head."""

    @staticmethod
    def forward(ctx, model, images, anchor):
        eng = model._engine
        z, _ = eng.forward(images, train=True, features_only=True)
        eng.generation += 1
        ctx.model, ctx.B, ctx.gen = model, images.shape[0], eng.generation
        return z

    @staticmethod
    def backward(ctx, dz):
        model = ctx.model
        eng = model._engine
        if ctx.gen != eng.generation:
            raise RuntimeError("backward through a stale forward: the activation workspace holds a later forward "
                               "(run forward/backward pairs in order)")
        model._bind_grads()
        eng.backward_features(ctx.B, dz.contiguous().float())
        if eng.fused_optimizer is None or eng.fused_optimizer() is None:
            eng.amp_update()
        return None, None, None


class VisionTransformerBase(_Base):
    """Same constructor surface as the reference base class (vision_transformer_base.py:288-402)."""

    def __init__(self, img_size: int = 256, patch_size: int = 16, in_chans: int = 1, num_classes: int = 2,
                 embed_dim: int = 768, depth: int = 12, num_heads: int = 12, mlp_ratio: float = 4.0, qkv_bias: bool = True,
                 representation_size: Optional[int] = None, distilled: bool = False, drop_rate: float = 0.0,
                 attn_drop_rate: float = 0.0, drop_path_rate: float = 0.0, embed_layer=None, norm_layer=None, act_layer=None,
                 weight_init: str = "", class_token: bool = True, no_embed_class: bool = False,
                 pos_embed_type: str = "learnable", pool_type: str = "cls", quality_aware: bool = True,
                 store_attention: bool = True, **kwargs):
        super().__init__()
        norm_layer = get_layer_from_string(norm_layer) if norm_layer else nn.LayerNorm
        act_layer = get_layer_from_string(act_layer) if act_layer else nn.GELU
        if embed_layer == "PatchEmbed" or embed_layer is None:
            embed_layer = PatchEmbed
        hp = dict(img_size=img_size, patch_size=patch_size, in_chans=in_chans, num_classes=num_classes, embed_dim=embed_dim,
                  depth=depth, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                  representation_size=representation_size, distilled=distilled, drop_rate=drop_rate,
                  attn_drop_rate=attn_drop_rate, drop_path_rate=drop_path_rate, embed_layer=embed_layer,
                  norm_layer=norm_layer, act_layer=act_layer, weight_init=weight_init, class_token=class_token,
                  no_embed_class=no_embed_class, pos_embed_type=pos_embed_type, pool_type=pool_type,
                  quality_aware=quality_aware, store_attention=store_attention)
        hp.update(kwargs)
        object.__setattr__(self, "_vitk_hparams", _HParams(hp))

        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.num_patches = (img_size // patch_size) ** 2
        self.class_token = class_token
        self.pool_type = pool_type
        self.store_attention = store_attention
        self.in_chans = in_chans
        self.patch_embed = embed_layer(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                       projection_type=kwargs.get("projection_type", "conv"), quality_aware=quality_aware)
        if class_token:
            self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        num_positions = self.num_patches + (1 if class_token else 0)
        if pos_embed_type == "learnable":
            self.pos_embed = nn.Parameter(torch.zeros(1, num_positions, embed_dim))
        elif pos_embed_type == "sinusoidal":
            self.register_buffer("pos_embed", self._create_sinusoidal_embedding(num_positions, embed_dim))
        else:
            raise ValueError(f"Unknown position embedding type: {pos_embed_type}")
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.blocks = None
        self.norm = norm_layer(embed_dim)
        if representation_size:
            self.pre_logits = nn.Sequential(nn.Linear(embed_dim, representation_size), nn.Tanh())
        else:
            self.pre_logits = nn.Identity()
        self.head = nn.Linear(self.num_features, num_classes)
        self._init_weights()
        self._attention_storage: List[torch.Tensor] = []
        self._engine: Optional[VitEngine] = None
        self._engine_key = None
        self._shadow_version = -1

    # Lightning exposes `hparams` as a property of its own; without Lightning we provide it.
    if _pl is None:
        @property
        def hparams(self):
            return self._vitk_hparams

        def log(self, *a, **k):
            pass

    def _create_sinusoidal_embedding(self, num_positions: int, embed_dim: int) -> torch.Tensor:
        """Fixed sin/cos table, interleaved (even columns sin, odd columns cos) -- vision_transformer_base.py:404-413."""
        inv_freq = torch.exp(torch.arange(0, embed_dim, 2) * (-math.log(10000.0) / embed_dim))
        angles = torch.outer(torch.arange(num_positions).float(), inv_freq)            # [positions, D/2]
        return torch.stack((angles.sin(), angles.cos()), dim=-1).reshape(1, num_positions, embed_dim)

    def _init_weights(self):
        """trunc_normal(0.02) linears, unit LayerNorms, He-style patch conv (vision_transformer_base.py:415-438)."""
        for _, m in self.named_modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        if hasattr(self.patch_embed, "proj") and isinstance(self.patch_embed.proj, nn.Conv2d):
            fan_in = self.patch_embed.proj.in_channels * self.patch_embed.proj.kernel_size[0] ** 2
            nn.init.trunc_normal_(self.patch_embed.proj.weight, std=math.sqrt(2.0 / fan_in))
        if isinstance(getattr(self, "pos_embed", None), nn.Parameter):
            nn.init.trunc_normal_(self.pos_embed, std=0.02)
        if hasattr(self, "cls_token"):
            nn.init.trunc_normal_(self.cls_token, std=0.02)

    # ------------------------------------------------------------------ engine plumbing
    def _n_prefix(self) -> int:
        return 1 if self.class_token else 0

    def _n_out(self) -> int:
        return 1

    def _pool_range(self):
        """Token range averaged after the final norm, or None for class-token pooling (vision_transformer_base.py:470-474:
        'cls' needs a class token; anything else is the mean of x[:, 1:], or of every token without a class token)."""
        if self.pool_type == "cls" and self.class_token:
            return None
        n = self.num_patches + self._n_prefix()
        return (1, n) if self.class_token else (0, n)

    def _ln_eps(self) -> float:
        """One epsilon for every LayerNorm of the model (norm_layer is a single factory in the reference, :263,273,377)."""
        norms = [self.norm] + [n for b in self.blocks for n in (b.norm1, b.norm2)]
        bad = [type(n).__name__ for n in norms if not isinstance(n, nn.LayerNorm)]
        if bad:
            raise NotImplementedError(f"the sm_100a path implements nn.LayerNorm norm layers only (got {sorted(set(bad))})")
        eps = {float(n.eps) for n in norms}
        if len(eps) != 1:
            raise NotImplementedError(f"all LayerNorm layers must share one eps (got {sorted(eps)})")
        return eps.pop()

    def _rep_size(self) -> int:
        if isinstance(self.pre_logits, nn.Identity):
            return 0
        rep = self.pre_logits[0].out_features
        if rep != self.head.in_features:      # the reference's head is Linear(embed_dim, classes) whatever representation_size is
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied: pre_logits emits {rep} features, head expects "
                               f"{self.head.in_features} (the reference fails the same way unless representation_size == embed_dim)")
        return rep

    def _check_supported(self) -> None:
        pass        # every constructor option of the reference classes is served

    def _engine_params(self) -> "OrderedDict[str, nn.Parameter]":
        skip = ("patch_embed.quality_score",)
        return OrderedDict((n, p) for n, p in self.named_parameters() if not n.startswith(skip))

    def _engine_tensors(self) -> "OrderedDict[str, torch.Tensor]":
        """Everything that lives in the engine's flat buffer: the trainable parameters plus a fixed (sinusoidal)
        `pos_embed` buffer, which gets a slot but never an optimizer group (its learning rate stays 0)."""
        named = self._engine_params()
        pe = getattr(self, "pos_embed", None)
        if pe is not None and not isinstance(pe, nn.Parameter):
            named["pos_embed"] = pe
        return named

    def _ensure_engine(self) -> VitEngine:
        named = self._engine_tensors()
        first = next(iter(named.values()))
        _require_cuda(first, type(self).__name__)
        key = (first.device, tuple((n, p.data_ptr()) for n, p in list(named.items())[:3]))
        if self._engine is None or self._engine_key != key:
            blk = self.blocks[0]
            dims = Dims(img=self.patch_embed.img_size, patch=self.patch_embed.patch_size, chans=self.in_chans,
                        dim=self.embed_dim, depth=len(self.blocks), heads=blk.attn.num_heads,
                        hidden=blk.mlp.fc1.out_features, classes=self.num_classes, n_prefix=self._n_prefix(),
                        n_out=self._n_out(), pool=self._pool_range(), rep=self._rep_size(),
                        patch_linear=self.patch_embed.projection_type != "conv", eps=self._ln_eps())
            eng = VitEngine(dims, OrderedDict((n, p.data) for n, p in named.items()), first.device,
                            dtype16=self._vitk_hparams.get("compute_dtype", torch.float16))
            for n, p in named.items():        # re-point the module's parameters at the flat buffers
                p.data = eng.flat.view(eng.flat.params, n)
                if isinstance(p, nn.Parameter):
                    p.grad = None
            eng.set_drop_path([float(getattr(b.drop_path, "drop_prob", 0.0)) for b in self.blocks])
            eng.set_dropout(float(self.pos_drop.p))       # pos_drop, proj_drop and Mlp.drop all carry drop_rate
            eng.set_attn_dropout(float(self.blocks[0].attn.attn_drop.p))
            if not isinstance(self.pos_embed, nn.Parameter):
                eng.frozen.add("pos_embed")
            self._engine = eng
            named = self._engine_params()
            first = next(iter(named.values()))
            self._engine_key = (first.device, tuple((n, p.data_ptr()) for n, p in list(named.items())[:3]))
            self._shadow_version = self._param_version()
        return self._engine

    def _param_version(self) -> int:
        return sum(p._version for p in self._engine_tensors().values())

    def _sync_shadow(self) -> None:
        """Re-cast the bf16 tensor-core shadow when a parameter was modified through torch (optimizer.step(),
        load_state_dict, ...).  The fused AdamW updates the shadow itself and does not bump versions."""
        v = self._param_version()
        if v != self._shadow_version:
            self._engine.flat.refresh_shadow()
            self._shadow_version = v

    def _bind_grads(self) -> None:
        """Point every param.grad at its slice of the flat gradient buffer.  If ALL grads are None (fresh model or
        optimizer.zero_grad(set_to_none=True)) the flat buffer is cleared first; otherwise kernels accumulate."""
        eng = self._engine
        named = self._engine_params()
        if all(p.grad is None for p in named.values()):
            eng.zero_grad()
        for n, p in named.items():
            if p.grad is None:
                p.grad = eng.flat.view(eng.flat.grads, n)

    def zero_grad(self, set_to_none: bool = True) -> None:
        """Gradients live in the engine's flat buffer: one memset instead of per-tensor work."""
        if self._engine is not None:
            self._engine.zero_grad()
        for n, p in self.named_parameters():
            if n.startswith("patch_embed.quality_score"):
                p.grad = None

    def _logits(self, x: torch.Tensor):
        B, C, H, W = x.shape
        if self.patch_embed.strict_img_size:
            assert H == self.patch_embed.img_size and W == self.patch_embed.img_size, \
                f"Input size ({H}x{W}) doesn't match expected size ({self.patch_embed.img_size}x{self.patch_embed.img_size})"
        self._check_supported()
        eng = self._ensure_engine()
        _require_cuda(x, type(self).__name__)
        self._sync_shadow()
        if self.training and torch.is_grad_enabled():
            anchor = eng.flat.params.new_zeros((), requires_grad=True)
            return _EncoderFn.apply(self, x, anchor)
        probs = [] if (self.store_attention and not self.training) else None
        l0, l1 = eng.forward(x, train=False, attn_probs=probs)
        if probs is not None:
            self._attention_storage = probs
            for blk, pm in zip(self.blocks, probs):
                blk.attn.attention_maps = pm
        return (l0, l1) if l1 is not None else l0

    # ------------------------------------------------------------------ reference API
    def _features(self, x: torch.Tensor) -> dict:
        """Inference-only encoder pass that also returns the pre-head features (no autograd graph)."""
        if self.training and torch.is_grad_enabled():
            raise NotImplementedError("DeiT.forward_features (the normalised token sequence) and extract_features are "
                                      "inference-only in the sm_100a path (call them under torch.no_grad() or model.eval()); "
                                      "the base class's forward_features() and forward() train")
        self._check_supported()
        eng = self._ensure_engine()
        _require_cuda(x, type(self).__name__)
        self._sync_shadow()
        probs = [] if (self.store_attention and not self.training) else None
        feats: dict = {}
        eng.forward(x, train=False, attn_probs=probs, features=feats)
        if probs is not None:
            self._attention_storage = probs
            for blk, pm in zip(self.blocks, probs):
                blk.attn.attention_maps = pm
        return feats

    def forward_features(self, x: torch.Tensor):
        """vision_transformer_base.py:440-479: (norm(x)[:, 0] as [B, D], quality_scores).  The quality branch is dead
        code in the reference's forward (its scores never reach the tokens), so None is returned for it.  In training mode
        with gradients enabled the feature carries an autograd edge into the encoder (a custom head can be trained on top)."""
        if self.training and torch.is_grad_enabled() and type(self).forward_features is VisionTransformerBase.forward_features:
            B, C, H, W = x.shape
            if self.patch_embed.strict_img_size:
                assert H == self.patch_embed.img_size and W == self.patch_embed.img_size, \
                    f"Input size ({H}x{W}) doesn't match expected size ({self.patch_embed.img_size}x{self.patch_embed.img_size})"
            eng = self._ensure_engine()
            _require_cuda(x, type(self).__name__)
            self._sync_shadow()
            anchor = eng.flat.params.new_zeros((), requires_grad=True)
            return _FeaturesFn.apply(self, x, anchor), None
        return self._features(x)["pooled"][0], None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """vision_transformer_base.py:482-486."""
        return self._logits(x)

    def get_attention_maps(self) -> Optional[torch.Tensor]:
        """[depth, B, H, N, N] stacked eval-mode attention maps (vision_transformer_base.py:488-492)."""
        if not self._attention_storage:
            return None
        return torch.stack([m.detach().cpu() for m in self._attention_storage])

    def extract_features(self, x: torch.Tensor) -> torch.Tensor:
        """vision_transformer_base.py:494-497."""
        return self._features(x)["pooled"][0]

    def training_step(self, batch, batch_idx):
        from .training import fused_cross_entropy
        x, y = batch
        loss, stats = fused_cross_entropy(self(x), y)
        self.log("train_loss", loss, on_step=True, on_epoch=True, prog_bar=True)
        self.log("train_acc", stats["acc"], on_step=True, on_epoch=True, prog_bar=True)
        return loss

    def validation_step(self, batch, batch_idx):
        from .training import fused_cross_entropy
        x, y = batch
        loss, stats = fused_cross_entropy(self(x), y)
        self.log("val_loss", loss, on_step=False, on_epoch=True, prog_bar=True)
        self.log("val_acc", stats["acc"], on_step=False, on_epoch=True, prog_bar=True)
        return {"val_loss": loss, "val_acc": stats["acc"]}

    def test_step(self, batch, batch_idx):
        from .training import fused_cross_entropy
        x, y = batch
        loss, stats = fused_cross_entropy(self(x), y)
        self.log("test_loss", loss, on_step=False, on_epoch=True)
        self.log("test_acc", stats["acc"], on_step=False, on_epoch=True)
        return {"test_loss": loss, "test_acc": stats["acc"]}

    def configure_optimizers(self):
        """vision_transformer_base.py:555-567 (AdamW over all parameters), on the fused multi-tensor kernel."""
        from .optim import FusedAdamW
        hp = self._vitk_hparams
        return FusedAdamW(self, lr=hp.get("learning_rate", 1e-3), weight_decay=hp.get("weight_decay", 0.05))

    def get_parameter_groups(self, weight_decay: float = 0.05, layer_decay: float = 0.75):
        """Per-parameter groups with layer-wise lr_scale -- vision_transformer_base.py:569-631, INCLUDING its
        substring-match behaviour ('blocks.1' also matches 'blocks.10'/'blocks.11', :600-603) so that optimizer
        parity with the reference holds."""
        param_groups = []
        no_decay = ["bias", "norm", "cls_token", "pos_embed"]
        if getattr(self, "blocks", None) is not None:
            num_layers = len(self.blocks)
            layer_scales = OrderedDict((f"blocks.{i}", layer_decay ** (num_layers - i - 1)) for i in range(num_layers))
            for name, param in self.named_parameters():
                if not param.requires_grad:
                    continue
                wd = 0.0 if any(nd in name for nd in no_decay) else weight_decay
                scale = 1.0
                for layer_name, layer_scale in layer_scales.items():
                    if layer_name in name:
                        scale = layer_scale
                        break
                param_groups.append({"params": [param], "weight_decay": wd, "lr_scale": scale, "name": name})
        else:
            for name, param in self.named_parameters():
                if not param.requires_grad:
                    continue
                wd = 0.0 if any(nd in name for nd in no_decay) else weight_decay
                param_groups.append({"params": [param], "weight_decay": wd, "name": name})
        return param_groups


class VisionTransformer(VisionTransformerBase):
    """vit_models.py:20-106."""

    def __init__(self, img_size: int = 256, patch_size: int = 16, in_chans: int = 1, num_classes: int = 2,
                 embed_dim: int = 768, depth: int = 12, num_heads: int = 12, mlp_ratio: float = 4.0, qkv_bias: bool = True,
                 representation_size: Optional[int] = None, drop_rate: float = 0.0, attn_drop_rate: float = 0.0,
                 drop_path_rate: float = 0.0, embed_layer=None, norm_layer=None, act_layer=None, **kwargs):
        norm_layer = nn.LayerNorm if norm_layer is None else norm_layer
        act_layer = nn.GELU if act_layer is None else act_layer
        super().__init__(img_size=img_size, patch_size=patch_size, in_chans=in_chans, num_classes=num_classes,
                         embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                         drop_rate=drop_rate, attn_drop_rate=attn_drop_rate, drop_path_rate=drop_path_rate,
                         norm_layer=norm_layer, act_layer=act_layer, **kwargs)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        norm_layer = get_layer_from_string(norm_layer)
        act_layer = get_layer_from_string(act_layer)
        self.blocks = nn.Sequential(*[
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=drop_rate,
                  attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer, act_layer=act_layer)
            for i in range(depth)])
        if representation_size and representation_size > 0:
            self.representation_size = representation_size
            self.pre_logits = nn.Sequential(nn.Linear(embed_dim, representation_size), nn.Tanh())
        else:
            self.representation_size = None
            self.pre_logits = nn.Identity()
        # NOTE (reference behaviour kept): _init_weights() ran in the base constructor BEFORE the blocks existed
        # (vision_transformer_base.py:402 vs vit_models.py:82), so block Linears keep nn.Linear's default init.


class ViTTiny(VisionTransformer):
    def __init__(self, **kwargs):
        super().__init__(embed_dim=192, depth=12, num_heads=3, mlp_ratio=4, **kwargs)


class ViTSmall(VisionTransformer):
    def __init__(self, **kwargs):
        super().__init__(embed_dim=384, depth=12, num_heads=6, mlp_ratio=4, **kwargs)


class ViTBase(VisionTransformer):
    def __init__(self, **kwargs):
        super().__init__(embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, **kwargs)


def create_vit_tiny(img_size: int = 256, patch_size: int = 16, in_chans: int = 1, num_classes: int = 2,
                    drop_path_rate: float = 0.1, **kwargs) -> ViTTiny:
    return ViTTiny(img_size=img_size, patch_size=patch_size, in_chans=in_chans, num_classes=num_classes,
                   drop_path_rate=drop_path_rate, **kwargs)


def create_vit_small(img_size: int = 256, patch_size: int = 16, in_chans: int = 1, num_classes: int = 2,
                     drop_path_rate: float = 0.1, **kwargs) -> ViTSmall:
    return ViTSmall(img_size=img_size, patch_size=patch_size, in_chans=in_chans, num_classes=num_classes,
                    drop_path_rate=drop_path_rate, **kwargs)


def create_vit_base(img_size: int = 256, patch_size: int = 16, in_chans: int = 1, num_classes: int = 2,
                    drop_path_rate: float = 0.1, **kwargs) -> ViTBase:
    return ViTBase(img_size=img_size, patch_size=patch_size, in_chans=in_chans, num_classes=num_classes,
                   drop_path_rate=drop_path_rate, **kwargs)


def create_vit_model(model_name: str, **kwargs) -> VisionTransformer:
    model_map = {"vit_tiny": create_vit_tiny, "vit_small": create_vit_small, "vit_base": create_vit_base}
    if model_name not in model_map:
        raise ValueError(f"Unknown ViT model: {model_name}. Available: {list(model_map.keys())}")
    return model_map[model_name](**kwargs)


# names the reference's tests and scripts import but the reference never defined
VIT_MODEL_REGISTRY = {"vit_tiny": create_vit_tiny, "vit_small": create_vit_small, "vit_base": create_vit_base}


def get_vit_model(model_name: str, **kwargs) -> VisionTransformer:
    if model_name not in VIT_MODEL_REGISTRY:
        raise ValueError(f"Unknown ViT model: {model_name}. Available: {list(VIT_MODEL_REGISTRY)}")
    return VIT_MODEL_REGISTRY[model_name](**kwargs)


VIT_PARAMS = {"vit_tiny": "5.7M", "vit_small": "22M", "vit_base": "86M"}


class DeiT(VisionTransformer):
    """deit_models.py:19-238: ViT + distillation token / head."""

    def __init__(self, img_size: int = 256, patch_size: int = 16, in_chans: int = 1, num_classes: int = 2,
                 embed_dim: int = 768, depth: int = 12, num_heads: int = 12, mlp_ratio: float = 4.0, qkv_bias: bool = True,
                 representation_size: Optional[int] = None, distilled: bool = True, drop_rate: float = 0.0,
                 attn_drop_rate: float = 0.0, drop_path_rate: float = 0.0, embed_layer=None, norm_layer=None, act_layer=None,
                 pretrained: bool = False, pretrained_cfg: Optional[Dict[str, Any]] = None, **kwargs):
        super().__init__(img_size=img_size, patch_size=patch_size, in_chans=in_chans, num_classes=num_classes,
                         embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                         representation_size=representation_size, drop_rate=drop_rate, attn_drop_rate=attn_drop_rate,
                         drop_path_rate=drop_path_rate, embed_layer=embed_layer, norm_layer=norm_layer, act_layer=act_layer,
                         **kwargs)
        self.distilled = distilled
        self.pretrained = pretrained
        self.pretrained_cfg = pretrained_cfg or {}
        self._vitk_hparams["distilled"] = distilled
        if self.distilled:
            self.dist_token = nn.Parameter(torch.zeros(1, 1, self.embed_dim))
            self.num_tokens = 2
            self.head_dist = nn.Linear(self.embed_dim, num_classes)
            if self.representation_size:
                # deit_models.py:84-99 replaces head_dist by a Sequential and then initialises `.weight` of it
                raise AttributeError("'Sequential' object has no attribute 'weight' (the reference's distilled DeiT cannot be "
                                     "built with representation_size, deit_models.py:84-99)")
            self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + self.num_tokens, self.embed_dim))
            nn.init.trunc_normal_(self.dist_token, std=0.02)
            nn.init.trunc_normal_(self.head_dist.weight, std=0.02)
            nn.init.zeros_(self.head_dist.bias)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        if pretrained:
            self.load_pretrained_weights()

    def _n_prefix(self) -> int:
        return 2 if self.distilled else 1

    def _n_out(self) -> int:
        return 2 if self.distilled else 1

    def _pool_range(self):
        return None          # DeiT.forward reads x[:, 0] / x[:, 1] whatever pool_type says (deit_models.py:220-238)

    def load_pretrained_weights(self, source=None):
        """deit_models.py:109-139.  The reference pulls ImageNet weights through timm + the network; here `source` (or
        `pretrained_cfg['file']`) names a local checkpoint / state_dict, adapted by the same three rules (pretrained.py)."""
        from . import pretrained
        return pretrained.load_into(self, source)

    def _adapt_pretrained_weights(self, state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """deit_models.py:141-164."""
        from . import pretrained
        return pretrained.adapt_state_dict(self, state_dict)

    def _interpolate_pos_embed(self, pos_embed: torch.Tensor) -> torch.Tensor:
        """deit_models.py:166-188.  (The reference reads `self.num_tokens`, which only a distilled model defines; a
        non-distilled model keeps its single class-token row here instead of raising AttributeError.)"""
        from . import pretrained
        return pretrained.resize_position_table(pos_embed, self._n_prefix(), self.patch_embed.num_patches)

    def get_attention_maps(self):
        """The reference's DeiT override forgets to fill the storage and returns None;
        per-block `attn.attention_maps` is the documented access path and is populated.  We return the stack."""
        return super().get_attention_maps()

    def forward_features(self, x: torch.Tensor) -> torch.Tensor:
        """deit_models.py:190-218: the normalised token sequence [B, T, D] (inference-only here)."""
        x_last = self._features(x)["x_last"]
        B, T, D = x_last.shape
        y, _, _ = ops.layernorm_fwd(x_last.reshape(B * T, D), self.norm.weight.detach(), self.norm.bias.detach(), eps=self.norm.eps,
                                    dtype=self._engine.dt16)
        return y.float().view(B, T, D)       # 16-bit rounding of the normalised tokens (operand precision of the path)

    def forward(self, x: torch.Tensor):
        """deit_models.py:220-238: (cls, dist) logits in training, their mean in eval."""
        out = self._logits(x)
        if self.distilled:
            l0, l1 = out
            if self.training:
                return l0, l1
            return (l0 + l1) / 2
        return out


class DeiTTiny(DeiT):
    def __init__(self, **kwargs):
        for k, v in dict(embed_dim=192, depth=12, num_heads=3, mlp_ratio=4).items():
            kwargs.setdefault(k, v)
        super().__init__(**kwargs)


class DeiTSmall(DeiT):
    def __init__(self, **kwargs):
        for k, v in dict(embed_dim=384, depth=12, num_heads=6, mlp_ratio=4).items():
            kwargs.setdefault(k, v)
        super().__init__(**kwargs)


class DeiTBase(DeiT):
    def __init__(self, **kwargs):
        for k, v in dict(embed_dim=768, depth=12, num_heads=12, mlp_ratio=4).items():
            kwargs.setdefault(k, v)
        super().__init__(**kwargs)


def _deit_factory(cls, timm_name):
    def create(img_size: int = 256, patch_size: int = 16, in_chans: int = 1, num_classes: int = 2, distilled: bool = True,
               pretrained: bool = False, **kwargs):
        if "pretrained_cfg" not in kwargs and pretrained:
            kwargs["pretrained_cfg"] = {"model_name": timm_name, "num_classes": 1000, "input_size": [3, 224, 224]}
        return cls(img_size=img_size, patch_size=patch_size, in_chans=in_chans, num_classes=num_classes,
                   distilled=distilled, pretrained=pretrained, **kwargs)
    return create


create_deit_tiny = _deit_factory(DeiTTiny, "deit_tiny_patch16_224")
create_deit_small = _deit_factory(DeiTSmall, "deit_small_patch16_224")
create_deit_base = _deit_factory(DeiTBase, "deit_base_patch16_224")


def create_deit_model(model_name: str, **kwargs) -> DeiT:
    """deit_models.py:385-413."""
    model_map = {"deit_tiny": create_deit_tiny, "deit_small": create_deit_small, "deit_base": create_deit_base}
    if model_name not in model_map:
        raise ValueError(f"Unknown DeiT model: {model_name}. Available: {list(model_map.keys())}")
    pretrained_cfg = kwargs.pop("pretrained_cfg", None)
    if pretrained_cfg is not None:
        return model_map[model_name](pretrained_cfg=pretrained_cfg, **kwargs)
    return model_map[model_name](**kwargs)
