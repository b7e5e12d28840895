"""Host-side engine of the B200 ViT/DeiT path: owns the flat parameter / gradient buffers and the
activation workspace, and sequences libvitk kernels for forward and backward.

This is orchestration only -- every FLOP and every byte moved happens inside libvitk.so
(include/vitk.h).  The kernel sequence restates, op for op, the reference's
DeiT.forward / VisionTransformerBase.forward (deit_models.py:190-238,
vision_transformer_base.py:120-143,174-195,217-223,282-285,440-486) and their autograd backward.

Memory layout in HBM (per rank)
  flat_params   fp32 [P]   all trainable tensors, each padded to 128 elements, in REVERSE execution order
                            (heads, norm, blocks L-1..0, patch/pos/cls) so gradient buckets complete in order
  flat_grads    fp32 [P]   same offsets, TRUE (unscaled) gradients; kernels ACCUMULATE into it
                            (split-K red.add, column-sum atomics)
  flat_16       fp16 [P]   tensor-core shadow of flat_params (written by the AdamW kernel / cast kernel)
  residual x    fp32 [B*T, D] per block boundary (2L+1 buffers, saved for backward)
  activations   fp16: LN outputs, qkv [B,T,3,H,64], attention out [B,T,H,64], fc1 GELU output and GELU derivative [B*T,4D], patches
  amp_state     fp32 [8]   {S, 1/S, good_steps, skipped, overflowed, ...}: dynamic loss scale, device resident

Numerics.  tcgen05 kind::f16 cannot mix fp16 with bf16 operands in one MMA (probed on B200: illegal
instruction) and pure bf16 operands miss BASELINE.json's parity bounds (measured: logits 1.1e-3 rms / 2.8e-3
max at B=32, gradients 0.6-1.7 % rel-L2).  So every tensor-core operand is fp16 (11-bit significand), with
fp32 accumulation, fp32 LN/softmax statistics, an fp32 residual stream and fp32 master weights; activation
GRADIENTS are stored as fp16 times a dynamic loss scale S (GradScaler semantics, bookkeeping on the device),
and every kernel that emits a parameter gradient multiplies by 1/S on the fly.  `dtype=torch.bfloat16` is
kept as a constructor option (S fixed at 1).
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import os
import torch

from . import _lib, ops

PAD = 128  # every tensor starts on a 128-element boundary (512 B fp32 / 256 B 16-bit: TMA + 128-bit safe)


@dataclass
class Dims:
    img: int
    patch: int
    chans: int
    dim: int
    depth: int
    heads: int
    hidden: int
    classes: int
    n_prefix: int      # 1 (cls) or 2 (cls + dist)
    n_out: int         # number of classification heads (2 for a distilled DeiT)
    pool: Optional[Tuple[int, int]] = None   # token range [t0, t1) averaged after the final norm ('gap' / no class token);
                                             # None = the fused class-token tail (vitk_head_fwd)
    rep: int = 0       # representation_size of `pre_logits` = Linear + Tanh (0 = Identity)
    patch_linear: bool = False   # PatchEmbed(projection_type='linear'): channel-last patch vectors, proj = Sequential[1]
    eps: float = 1e-5  # LayerNorm epsilon of every norm layer (nn.LayerNorm default; a partial(nn.LayerNorm, eps=...) changes it)

    @property
    def n_patches(self) -> int:
        return (self.img // self.patch) ** 2

    @property
    def tokens(self) -> int:
        return self.n_patches + self.n_prefix

    @property
    def kpatch(self) -> int:
        return self.chans * self.patch * self.patch


def execution_order(names: List[str], depth: int) -> List[str]:
    """Reverse execution order: the order in which backward finishes each tensor's gradient."""
    def key(n: str):
        if n.startswith("head") or n.startswith("norm.") or n.startswith("pre_logits."):
            return (0, 0)
        if n.startswith("blocks."):
            return (1, depth - 1 - int(n.split(".")[1]))
        return (2, 0)
    return sorted(names, key=lambda n: (key(n), names.index(n)))


@dataclass
class GraySpec:
    """How single-channel tiles become the model's `chans` input channels: optional per-image clamp-normalise bounds
    (fp32 CUDA [B,2], ingest percentile stage) and optional T.Normalize statistics (one per channel)."""
    bounds: Optional[torch.Tensor] = None
    mean: Optional[Tuple[float, ...]] = None
    std: Optional[Tuple[float, ...]] = None


class FlatParams:
    """One fp32 buffer for all trainable tensors + same-layout gradient and 16-bit shadow buffers."""

    def __init__(self, named: "OrderedDict[str, torch.Tensor]", depth: int, device, dtype16=torch.float16):
        order = execution_order(list(named), depth)
        self.offsets: Dict[str, Tuple[int, torch.Size]] = {}
        total = 0
        for n in order:
            self.offsets[n] = (total, named[n].shape)
            total += (named[n].numel() + PAD - 1) // PAD * PAD
        self.numel = total
        self.order = order
        self.dtype16 = dtype16
        self.params = torch.zeros(total, dtype=torch.float32, device=device)
        self.grads = torch.zeros(total, dtype=torch.float32, device=device)
        self.w16 = torch.zeros(total, dtype=dtype16, device=device)
        for n in order:
            self.view(self.params, n).copy_(named[n].detach().to(device=device, dtype=torch.float32))

    def view(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        off, shape = self.offsets[name]
        return buf[off:off + shape.numel()].view(shape)

    def refresh_shadow(self) -> None:
        """flat_params -> 16-bit shadow (one vectorised cast kernel over the whole buffer)."""
        if self.dtype16 == torch.float16:
            ops.cast_fp16(self.params, self.w16)
        else:
            ops.cast_bf16(self.params, self.w16)

    def bucket_slices(self, bucket_bytes: int) -> List[Tuple[int, int]]:
        """Contiguous [start, end) element ranges of ~bucket_bytes, cut at tensor boundaries, in
        gradient-completion order (used by the data-parallel all-reduce)."""
        out, start, cur = [], 0, 0
        for n in self.order:
            off, shape = self.offsets[n]
            end = off + (shape.numel() + PAD - 1) // PAD * PAD
            cur = end
            if (cur - start) * 4 >= bucket_bytes:
                out.append((start, cur))
                start = cur
        if cur > start:
            out.append((start, cur))
        return out


class Workspace:
    """Activation + scratch buffers for one (batch, training?) shape; allocated once, reused every step."""

    def __init__(self, d: Dims, B: int, train: bool, device, dt16):
        self.B, self.train = B, train
        T, D, H = d.tokens, d.dim, d.heads
        M = B * T
        f32 = torch.float32
        e = lambda *s, dt=dt16: torch.empty(*s, dtype=dt, device=device)
        L = d.depth if train else 1
        self.patches = e(B * d.n_patches, d.kpatch)
        nres = 2 * d.depth + 1 if train else 3
        self.x = [e(M, D, dt=f32) for _ in range(nres)]
        self.xn1 = [e(M, D) for _ in range(L)]
        self.xn2 = [e(M, D) for _ in range(L)]
        self.stats = [e(4, M, dt=f32) for _ in range(L)]          # mean1, rstd1, mean2, rstd2
        self.qkv = [e(M, 3 * D) for _ in range(L)]
        self.ao = [e(M, D) for _ in range(L)]
        self.lse = [e(B, H, T, dt=f32) for _ in range(L)]
        self.dact = [e(M, d.hidden) for _ in range(L)]
        self.act = [e(M, d.hidden) for _ in range(L)]
        if train:
            self.dx = [e(M, D, dt=f32) for _ in range(2)]
            self.dx16 = e(M, D)
            self.dxn = e(M, D)
            self.d_ao = e(M, D)
            self.d_pre = e(M, d.hidden)
            self.dqkv = e(M, 3 * D)
            self.delta = e(B, H, T, dt=f32)
            self.dpatch = e(B * d.n_patches, D)
        self.head_saved = None
        # the last block's attn.proj / norm2 / Mlp on the rows the classifier reads (tokens 0..n_out-1 of every image)
        R = B * d.n_out
        self.c_ao, self.c_xin, self.c_xmid, self.c_xout = e(R, D), e(R, D, dt=f32), e(R, D, dt=f32), e(R, D, dt=f32)
        self.c_xn2, self.c_stats = e(R, D), e(2, R, dt=f32)
        self.c_dact, self.c_act = e(R, d.hidden), e(R, d.hidden)
        self.pruned = False
        if train:
            self.c_dx = [e(R, D, dt=f32) for _ in range(2)]
            self.c_dx16, self.c_dxn, self.c_dao, self.c_dpre = e(R, D), e(R, D), e(R, D), e(R, d.hidden)

    def nbytes(self) -> int:
        tot = 0
        for v in self.__dict__.values():
            for t in (v if isinstance(v, list) else [v]):
                if isinstance(t, torch.Tensor):
                    tot += t.numel() * t.element_size()
        return tot


class VitEngine:
    """Forward / backward of the whole encoder for one model instance."""

    INIT_LOSS_SCALE = 65536.0
    GROWTH_INTERVAL = 2000

    def __init__(self, dims: Dims, named_params: "OrderedDict[str, torch.Tensor]", device, dtype16=torch.float16):
        if dims.dim % dims.heads != 0 or dims.dim // dims.heads != 64:
            raise NotImplementedError(
                f"libvitk attention kernels cover head_dim 64 (got dim={dims.dim}, heads={dims.heads}); "
                "every ViT/DeiT tiny/small/base variant of the reference satisfies this")
        if dims.dim % 8 or dims.hidden % 8 or dims.kpatch % 8 or dims.patch % 8:
            raise NotImplementedError("embed_dim, mlp hidden and patch size must be multiples of 8")
        _lib.load()
        self.d = dims
        self.device = torch.device(device)
        self.dt16 = dtype16
        self.flat = FlatParams(named_params, dims.depth, self.device, dtype16)
        self.flat.refresh_shadow()
        self._ws: Dict[Tuple[int, bool], Workspace] = {}
        self.scale = 64 ** -0.5
        self.sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        self.grad_ready_hook = None   # callable(stage) used by the data-parallel bucket launcher
        self.fused_optimizer = None   # weakref to the FusedAdamW that owns the loss-scale bookkeeping (None: external optimizer)
        self.generation = 0
        self._drop_on = False
        s0 = self.INIT_LOSS_SCALE if dtype16 == torch.float16 else 1.0
        self.amp = torch.tensor([s0, 1.0 / s0, 0, 0, 0, 0, 0, 0], dtype=torch.float32, device=self.device)
        self.amp_scratch = torch.zeros(4, dtype=torch.float32, device=self.device)
        self.growth_interval = self.GROWTH_INTERVAL if dtype16 == torch.float16 else 0
        # stochastic depth (DropPath, vision_transformer_base.py:56-64, rates linspace(0, dpr, L) per block, vit_models.py:73):
        # branch 2l = attention branch of block l, branch 2l+1 = its MLP branch; one Bernoulli draw per (branch, sample)
        self.drop_path = [0.0] * (2 * dims.depth)
        self._dp_prob_dev = None
        self.last_drop_scale = None     # [2L, B] factors of the latest training forward (tests / debugging)
        # nn.Dropout(drop_rate) sites (pos_drop, proj_drop, Mlp.drop x2): counter-based masks recomputed from one device
        # seed per step (include/vitk.h `vitk_dropout`); site 0 = pos_drop, block l: 1+3l proj, 2+3l after GELU, 3+3l after fc2
        self.drop_rate = 0.0
        self.attn_drop_rate = 0.0       # Attention.attn_drop (:184): site ATTN_SITE0 + l, element ((b*H + h)*T + q)*Tpad + key
        self.drop_seed = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.frozen = set()             # flat-buffer entries that are not trained (a sinusoidal pos_embed buffer)
        # The classifier reads x[:, 0] (and x[:, 1] for the distillation head) of the last block's output and nothing else
        # (vision_transformer_base.py:474-479, deit_models.py:224-235), so with class-token pooling the last block's attn.proj,
        # norm2 and Mlp -- forward and backward -- only matter on those rows: they run on B*n_out rows instead of B*T.  Logits,
        # loss and every parameter gradient are unchanged (the other rows' outputs are never read, their gradients are exactly
        # zero).  Off whenever something does read the other rows or indexes them: token pooling ('gap'), pre_logits tails,
        # forward_features / `features`, dropout and stochastic depth (their masks are indexed by dense row).
        self.cls_rows_last_block = os.environ.get("VITK_DENSE_LAST_BLOCK", "0") != "1"
        proj = "patch_embed.proj.1." if dims.patch_linear else "patch_embed.proj."
        self.patch_w, self.patch_b = proj + "weight", proj + "bias"

    def set_drop_path(self, per_block_rates) -> None:
        rates = [float(r) for r in per_block_rates]
        if len(rates) != self.d.depth or any(not (0.0 <= r < 1.0) for r in rates):
            raise ValueError("set_drop_path: one rate in [0, 1) per block expected")
        self.drop_path = [r for r in rates for _ in range(2)]
        self._dp_prob_dev = torch.tensor(self.drop_path, dtype=torch.float32, device=self.device)

    def set_dropout(self, rate: float) -> None:
        if not (0.0 <= float(rate) < 1.0):
            raise ValueError("set_dropout: rate in [0, 1) expected")
        self.drop_rate = float(rate)

    ATTN_SITE0 = 1000

    def set_attn_dropout(self, rate: float) -> None:
        if not (0.0 <= float(rate) < 1.0):
            raise ValueError("set_attn_dropout: rate in [0, 1) expected")
        self.attn_drop_rate = float(rate)

    def _attn_site(self, l: int):
        return (self.drop_seed, self.attn_drop_rate, self.ATTN_SITE0 + l) if (self._drop_on and self.attn_drop_rate > 0.0) else None

    def _site(self, site: int):
        """(seed, p, site) of one Dropout call, or None when dropout is off for this pass."""
        return (self.drop_seed, self.drop_rate, site) if (self._drop_on and self.drop_rate > 0.0) else None

    # ------------------------------------------------------------------ helpers
    def w(self, name: str) -> torch.Tensor:      # 16-bit shadow view
        return self.flat.view(self.flat.w16, name)

    def p(self, name: str) -> torch.Tensor:      # fp32 master view
        return self.flat.view(self.flat.params, name)

    def g(self, name: str) -> torch.Tensor:      # fp32 gradient view
        return self.flat.view(self.flat.grads, name)

    @property
    def loss_scale(self) -> torch.Tensor:        # device scalar S
        return self.amp[0:1]

    @property
    def grad_unscale(self) -> torch.Tensor:      # device scalar 1/S
        return self.amp[1:2]

    def workspace(self, B: int, train: bool) -> Workspace:
        key = (B, train)
        ws = self._ws.get(key)
        if ws is None:
            ws = Workspace(self.d, B, train, self.device, self.dt16)
            self._ws[key] = ws
        return ws

    def _wgrad(self, dy: torch.Tensor, x: torch.Tensor, wname: str, rows: int, bias_name: Optional[str] = None) -> None:
        """grad[wname][N_out, K_in] += (1/S) * dy[rows, N_out]^T @ x[rows, K_in] (both operands read MN-major);
        with `bias_name`, grad[bias][N_out] += (1/S) * column sums of dy from the same pass over dy."""
        gw = self.g(wname)
        n_out = gw.shape[0]
        k_in = gw.numel() // n_out
        ops.gemm(dy, x, n_out, k_in, rows, a_mn=True, b_mn=True, out=gw, epilogue=_lib.EPI_ATOMIC_ADD,
                 split_k=0, alpha_dev=self.grad_unscale,       # 0: the library picks the K split (tile shape / CTA pairs)
                 colsum_out=self.g(bias_name) if bias_name is not None else None)

    # ------------------------------------------------------------------ forward
    def forward(self, images: torch.Tensor, train: bool, attn_probs=None, features: Optional[dict] = None,
                gray: Optional["GraySpec"] = None, features_only: bool = False):
        """gray: when given, `images` are single-channel tiles [B,H,W] / [B,1,H,W] (fp32 in [0,1], raw uint16, fp16 or
        bf16) that the loader would have replicated to `chans` channels and normalised (vit_transforms.py:381-393).
        attn_probs (inference only): a list that receives one fp32 [B,H,T,T] tensor per block, or a preallocated
        contiguous fp32 [L,B,H,T,T] tensor the blocks write into.
        features (inference only): a dict that receives 'pooled' = norm(x)[:, :n_out] as fp32 [n_out,B,D] and
        'x_last' = the final residual stream fp32 [B,T,D] (a workspace view: copy or consume before the next call)."""
        d = self.d
        if gray is None and (images.dim() != 4 or images.shape[1] != d.chans):
            raise ValueError(f"expected images [B,{d.chans},H,W], got {tuple(images.shape)}")
        if gray is not None:
            if d.patch_linear:
                raise NotImplementedError("gray tile input is implemented for projection_type='conv' only")
            if images.dim() == 4 and images.shape[1] == 1:
                images = images[:, 0]
            if images.dim() != 3 or images.shape[1] != d.img or images.shape[2] != d.img:
                raise ValueError(f"expected gray tiles [B,{d.img},{d.img}] (or [B,1,H,W]), got {tuple(images.shape)}")
        B = images.shape[0]
        T, D = d.tokens, d.dim
        M = B * T
        ws = self.workspace(B, train)
        images = images.contiguous()
        if gray is None and images.dtype != torch.float32:
            images = images.float()
        dp = None
        if train and any(r > 0.0 for r in self.drop_path):
            # uniforms come from torch's CUDA generator (seedable, CUDA-graph safe); the mask arithmetic is libvitk's
            u = torch.rand(2 * d.depth, B, dtype=torch.float32, device=self.device)
            if getattr(ws, "dp", None) is None:
                ws.dp = torch.empty(2 * d.depth, M, dtype=torch.float32, device=self.device)
            ops.droppath_scale(u, self._dp_prob_dev, T, out=ws.dp)
            dp = ws.dp
            self.last_drop_scale = ws.dp[:, ::T]
        ws.dp_active = dp is not None
        self._drop_on = bool(train and (self.drop_rate > 0.0 or self.attn_drop_rate > 0.0))
        ws.drop_on = self._drop_on
        if self._drop_on:
            self.drop_seed.random_(0, 1 << 62)      # torch's CUDA generator: seedable and CUDA-graph safe (fresh per replay)
        rs = lambda i: dp[i] if (dp is not None and self.drop_path[i] > 0.0) else None
        if gray is not None:
            # single-channel tiles (fp32 | uint16 | 16-bit) -> replicated + normalised channels, written straight into the
            # 16-bit patch matrix: the fp32 [B,C,H,W] batch of the reference's loader never exists on this path
            ops.tiles_to_patches(images, d.chans, d.patch, bounds=gray.bounds, mean=gray.mean, std=gray.std, out=ws.patches)
        else:
            ops.patchify(images, d.patch, out=ws.patches, channel_last=d.patch_linear)
        x0 = ws.x[0]
        ops.gemm(ws.patches, self.w(self.patch_w), B * d.n_patches, D, d.kpatch, out=x0,
                 bias=self.p(self.patch_b), epilogue=_lib.EPI_TOKENS,
                 tokens=(d.n_patches, T, d.n_prefix), pos=self.p("pos_embed"), drop=self._site(0))
        ops.prefix_tokens_fwd(x0.view(B, T, D), self.p("cls_token") if d.n_prefix >= 1 else None,
                              self.p("dist_token") if d.n_prefix == 2 else None, self.p("pos_embed"), d.n_prefix,
                              drop=self._site(0))
        prune = (self.cls_rows_last_block and d.pool is None and not d.rep and features is None and not features_only
                 and dp is None and not self._drop_on and d.n_prefix >= d.n_out)
        ws.pruned = prune
        nr = d.n_out
        R = B * nr
        for l in range(d.depth):
            s = l if train else 0
            pre = f"blocks.{l}."
            if train:
                x_in, x_mid, x_out = ws.x[2 * l], ws.x[2 * l + 1], ws.x[2 * l + 2]
            else:
                x_in, x_mid, x_out = ws.x[(2 * l) % 3], ws.x[(2 * l + 1) % 3], ws.x[(2 * l + 2) % 3]
            st = ws.stats[s]
            ops.layernorm_fwd(x_in, self.p(pre + "norm1.weight"), self.p(pre + "norm1.bias"), eps=d.eps, y=ws.xn1[s], mean=st[0], rstd=st[1])
            ops.gemm(ws.xn1[s], self.w(pre + "attn.qkv.weight"), M, 3 * D, D, out=ws.qkv[s], bias=self.p(pre + "attn.qkv.bias"))
            probs = None
            probs_im = None
            if isinstance(attn_probs, torch.Tensor):       # caller-owned [L,B,H,T,T] buffer: no stacking copy later
                probs = attn_probs[l]
            elif isinstance(attn_probs, tuple):            # ("image_major", [B,L,H,T,T]): one image's maps contiguous (ensemble rollout)
                probs_im = attn_probs[1]
            elif attn_probs is not None:
                probs = torch.empty(B, d.heads, T, T, dtype=torch.float32, device=self.device)
                attn_probs.append(probs)
            # last block of a class-token model: only the classifier's query rows are consumed (and the maps need every lse)
            q_rows = nr if (prune and l == d.depth - 1 and attn_probs is None) else 0
            ops.attention_fwd(ws.qkv[s], B, T, d.heads, self.scale, out=ws.ao[s], lse=ws.lse[s], probs=probs, drop=self._attn_site(l),
                              q_rows=q_rows)
            if probs_im is not None:
                ops.attention_probs(ws.qkv[s], ws.lse[s], B, T, d.heads, self.scale, probs_im[:, l],
                                    batch_stride=d.depth * d.heads * T * T)
            if prune and l == d.depth - 1:
                ops.gather_rows(ws.ao[s].view(B, T, D), nr, ws.c_ao)
                ops.gather_rows(x_in.view(B, T, D), nr, ws.c_xin)
                ops.gemm(ws.c_ao, self.w(pre + "attn.proj.weight"), R, D, D, out=ws.c_xmid, bias=self.p(pre + "attn.proj.bias"),
                         residual=ws.c_xin)
                ops.layernorm_fwd(ws.c_xmid, self.p(pre + "norm2.weight"), self.p(pre + "norm2.bias"), eps=d.eps, y=ws.c_xn2,
                                  mean=ws.c_stats[0], rstd=ws.c_stats[1])
                ops.gemm(ws.c_xn2, self.w(pre + "mlp.fc1.weight"), R, d.hidden, D, out=ws.c_dact, out2=ws.c_act,
                         bias=self.p(pre + "mlp.fc1.bias"), epilogue=_lib.EPI_GELU)
                ops.gemm(ws.c_act, self.w(pre + "mlp.fc2.weight"), R, D, d.hidden, out=ws.c_xout, bias=self.p(pre + "mlp.fc2.bias"),
                         residual=ws.c_xmid)
                continue
            ops.gemm(ws.ao[s], self.w(pre + "attn.proj.weight"), M, D, D, out=x_mid, bias=self.p(pre + "attn.proj.bias"), residual=x_in,
                     row_scale=rs(2 * l), drop=self._site(1 + 3 * l))
            ops.layernorm_fwd(x_mid, self.p(pre + "norm2.weight"), self.p(pre + "norm2.bias"), eps=d.eps, y=ws.xn2[s], mean=st[2], rstd=st[3])
            ops.gemm(ws.xn2[s], self.w(pre + "mlp.fc1.weight"), M, d.hidden, D, out=ws.dact[s], out2=ws.act[s],
                     bias=self.p(pre + "mlp.fc1.bias"), epilogue=_lib.EPI_GELU, drop=self._site(2 + 3 * l))
            ops.gemm(ws.act[s], self.w(pre + "mlp.fc2.weight"), M, D, d.hidden, out=x_out, bias=self.p(pre + "mlp.fc2.bias"), residual=x_mid,
                     row_scale=rs(2 * l + 1), drop=self._site(3 + 3 * l))
        x_last = ws.x[2 * d.depth] if train else ws.x[(2 * d.depth) % 3]
        two = d.n_out == 2
        pooled = None
        if features is not None:
            pooled = torch.empty(d.n_out, B, D, dtype=torch.float32, device=self.device)
            features["pooled"] = pooled
            features["x_last"] = x_last.view(B, T, D)
        if features_only:
            # training-mode forward_features (vision_transformer_base.py:440-479): stop before the head; the saved state lets
            # backward_features() start from the gradient of the pre-head feature (norm -> pool -> pre_logits, fp32 tail kernels)
            return self._tail_fwd(ws, x_last.view(B, T, D), train, features, with_head=False), None
        if d.pool is not None or d.rep:
            return self._tail_fwd(ws, x_last.view(B, T, D), train, features), None
        l0, l1, xhat, rstd = ops.head_fwd(ws.c_xout.view(B, nr, D) if prune else x_last.view(B, T, D),
                                          self.p("norm.weight"), self.p("norm.bias"),
                                          self.p("head.weight"), self.p("head.bias"),
                                          self.p("head_dist.weight") if two else None,
                                          self.p("head_dist.bias") if two else None, d.n_out, eps=d.eps, pooled=pooled)
        if train:
            ws.head_saved = (xhat, rstd)
        return l0, l1

    # ------------------------------------------------------------------ general classification tail
    def _tail_range(self) -> Tuple[int, int]:
        return self.d.pool if self.d.pool is not None else (0, 1)

    def _tail_fwd(self, ws: Workspace, x_last: torch.Tensor, train: bool, features: Optional[dict], with_head: bool = True) -> torch.Tensor:
        """norm -> pool over a token range -> pre_logits -> head (vision_transformer_base.py:468-486) for the constructor
        options outside the fused class-token kernel: fp32 throughout, a few hundred KB per step."""
        t0, t1 = self._tail_range()
        pooled, mean, rstd = ops.pool_norm_fwd(x_last, self.p("norm.weight"), self.p("norm.bias"), t0, t1, eps=self.d.eps)
        z = pooled
        if self.d.rep:
            z = ops.dense_fwd(pooled, self.p("pre_logits.0.weight"), self.p("pre_logits.0.bias"), act=1)
        if features is not None:
            features["pooled"] = z.unsqueeze(0)       # forward_features returns the pre-head feature (after pre_logits)
        if train:
            ws.head_saved = (pooled, mean, rstd, z, x_last)
        if not with_head:
            return z
        return ops.dense_fwd(z, self.p("head.weight"), self.p("head.bias"), act=0)

    def _tail_bwd(self, ws: Workspace, dl0: torch.Tensor, dx: torch.Tensor, dcolsum, branch_scale, branch_drop,
                  from_features: bool = False) -> None:
        pooled, mean, rstd, z, x_last = ws.head_saved
        t0, t1 = self._tail_range()
        dz = dl0 if from_features else ops.dense_bwd(dl0, None, z, self.p("head.weight"), self.g("head.weight"), self.g("head.bias"), act=0)
        if self.d.rep:
            dz = ops.dense_bwd(dz, z, pooled, self.p("pre_logits.0.weight"), self.g("pre_logits.0.weight"),
                               self.g("pre_logits.0.bias"), act=1)
        ops.pool_norm_bwd(dz, x_last, mean, rstd, self.p("norm.weight"), dx.view_as(x_last), ws.dx16, self.g("norm.weight"),
                          self.g("norm.bias"), dcolsum, t0, t1, loss_scale=self.loss_scale, branch_scale=branch_scale,
                          branch_drop=branch_drop)

    # ------------------------------------------------------------------ backward
    def backward_features(self, B: int, dz: torch.Tensor) -> None:
        """Backward of a `features_only` training forward: dz = TRUE gradient of the pre-head feature [B, D | rep]."""
        self.backward(B, dz, None, from_features=True)

    def backward(self, B: int, dl0: torch.Tensor, dl1: Optional[torch.Tensor], from_features: bool = False) -> None:
        """Takes TRUE dlogits; accumulates every TRUE parameter gradient into flat.grads (no input gradient).
        Activation gradients in between are S-scaled 16-bit tensors."""
        d = self.d
        T, D = d.tokens, d.dim
        M = B * T
        ws = self.workspace(B, True)
        if ws.head_saved is None:
            raise RuntimeError("backward() called without a preceding training forward()")
        general_tail = d.pool is not None or bool(d.rep) or from_features
        two = d.n_out == 2
        u = self.grad_unscale
        dp = ws.dp if getattr(ws, "dp_active", False) else None
        self._drop_on = bool(getattr(ws, "drop_on", False))
        rs = lambda i: dp[i] if (dp is not None and i >= 0 and self.drop_path[i] > 0.0) else None
        dx, dx_alt = ws.dx[0], ws.dx[1]
        last_fc2_bias = self.g(f"blocks.{d.depth - 1}.mlp.fc2.bias")
        if general_tail:
            self._tail_bwd(ws, dl0.contiguous(), dx, last_fc2_bias, rs(2 * d.depth - 1), self._site(3 + 3 * (d.depth - 1)),
                           from_features=from_features)
        else:
            xhat, rstd = ws.head_saved
            pruned = bool(ws.pruned)
            ops.head_bwd(dl0.contiguous(), dl1.contiguous() if two else None, xhat, rstd, self.p("norm.weight"), self.p("norm.bias"),
                         self.p("head.weight"), self.p("head_dist.weight") if two else None,
                         ws.c_dx[0] if pruned else dx, ws.c_dx16 if pruned else ws.dx16,
                         self.g("norm.weight"), self.g("norm.bias"), self.g("head.weight"), self.g("head.bias"),
                         self.g("head_dist.weight") if two else None, self.g("head_dist.bias") if two else None,
                         last_fc2_bias, d.n_out if pruned else T, d.n_out, loss_scale=self.loss_scale,
                         branch_scale=rs(2 * d.depth - 1), branch_drop=self._site(3 + 3 * (d.depth - 1)))
        ws.head_saved = None
        self._notify("head")
        for l in range(d.depth - 1, -1, -1):
            pre = f"blocks.{l}."
            st = ws.stats[l]
            x_in, x_mid = ws.x[2 * l], ws.x[2 * l + 1]
            if l == d.depth - 1 and not general_tail and ws.pruned:
                # the last block's MLP and attn.proj on the classifier's rows only (see cls_rows_last_block); the gradient of
                # those rows then goes back into dense, otherwise-zero tensors for the attention backward and norm1
                nr = d.n_out
                R = B * nr
                self._wgrad(ws.c_dx16, ws.c_act, pre + "mlp.fc2.weight", R)
                ops.gemm(ws.c_dx16, self.w(pre + "mlp.fc2.weight"), R, d.hidden, D, b_mn=True, out=ws.c_dpre, aux=ws.c_dact,
                         epilogue=_lib.EPI_DGELU)
                self._wgrad(ws.c_dpre, ws.c_xn2, pre + "mlp.fc1.weight", R, bias_name=pre + "mlp.fc1.bias")
                ops.gemm(ws.c_dpre, self.w(pre + "mlp.fc1.weight"), R, D, d.hidden, b_mn=True, out=ws.c_dxn)
                ops.layernorm_bwd(ws.c_dxn, ws.c_xmid, ws.c_stats[0], ws.c_stats[1], self.p(pre + "norm2.weight"),
                                  self.g(pre + "norm2.weight"), self.g(pre + "norm2.bias"), dres=ws.c_dx[0], dx=ws.c_dx[1],
                                  dx16=ws.c_dx16, dcolsum=self.g(pre + "attn.proj.bias"), unscale=u)
                self._wgrad(ws.c_dx16, ws.c_ao, pre + "attn.proj.weight", R)
                ops.gemm(ws.c_dx16, self.w(pre + "attn.proj.weight"), R, D, D, b_mn=True, out=ws.c_dao)
                ops.expand_rows(ws.c_dao.view(B, nr, D), nr, ws.d_ao.view(B, T, D))
                ops.expand_rows(ws.c_dx[1].view(B, nr, D), nr, dx.view(B, T, D))
                ops.attention_bwd(ws.qkv[l], ws.ao[l], ws.d_ao, ws.lse[l], B, T, d.heads, self.scale, dqkv=ws.dqkv, delta=ws.delta,
                                  drop=self._attn_site(l), q_rows=nr)       # d_ao is zero from row nr on
                self._wgrad(ws.dqkv, ws.xn1[l], pre + "attn.qkv.weight", M, bias_name=pre + "attn.qkv.bias")
                ops.gemm(ws.dqkv, self.w(pre + "attn.qkv.weight"), M, D, 3 * D, b_mn=True, out=ws.dxn)
                prev_bias = self.g(f"blocks.{l - 1}.mlp.fc2.bias") if l > 0 else None
                ops.layernorm_bwd(ws.dxn, x_in, st[0], st[1], self.p(pre + "norm1.weight"), self.g(pre + "norm1.weight"),
                                  self.g(pre + "norm1.bias"), dres=dx, dx=dx_alt, dx16=ws.dx16 if l > 0 else None,
                                  dcolsum=prev_bias, unscale=u)
                dx, dx_alt = dx_alt, dx
                self._notify(pre)
                continue
            # ---- MLP branch: x_out = x_mid + fc2(gelu(fc1(norm2(x_mid))))
            self._wgrad(ws.dx16, ws.act[l], pre + "mlp.fc2.weight", M)
            ops.gemm(ws.dx16, self.w(pre + "mlp.fc2.weight"), M, d.hidden, D, b_mn=True, out=ws.d_pre, aux=ws.dact[l],
                     epilogue=_lib.EPI_DGELU)
            self._wgrad(ws.d_pre, ws.xn2[l], pre + "mlp.fc1.weight", M, bias_name=pre + "mlp.fc1.bias")
            ops.gemm(ws.d_pre, self.w(pre + "mlp.fc1.weight"), M, D, d.hidden, b_mn=True, out=ws.dxn)
            ops.layernorm_bwd(ws.dxn, x_mid, st[2], st[3], self.p(pre + "norm2.weight"), self.g(pre + "norm2.weight"),
                              self.g(pre + "norm2.bias"), dres=dx, dx=dx_alt, dx16=ws.dx16,
                              dcolsum=self.g(pre + "attn.proj.bias"), unscale=u, branch_scale=rs(2 * l),
                              branch_drop=self._site(1 + 3 * l))
            dx, dx_alt = dx_alt, dx
            # ---- attention branch: x_mid = x_in + proj(attn(qkv(norm1(x_in))))
            self._wgrad(ws.dx16, ws.ao[l], pre + "attn.proj.weight", M)
            ops.gemm(ws.dx16, self.w(pre + "attn.proj.weight"), M, D, D, b_mn=True, out=ws.d_ao)
            ops.attention_bwd(ws.qkv[l], ws.ao[l], ws.d_ao, ws.lse[l], B, T, d.heads, self.scale, dqkv=ws.dqkv, delta=ws.delta,
                              drop=self._attn_site(l))
            self._wgrad(ws.dqkv, ws.xn1[l], pre + "attn.qkv.weight", M, bias_name=pre + "attn.qkv.bias")
            ops.gemm(ws.dqkv, self.w(pre + "attn.qkv.weight"), M, D, 3 * D, b_mn=True, out=ws.dxn)
            prev_bias = self.g(f"blocks.{l - 1}.mlp.fc2.bias") if l > 0 else None
            ops.layernorm_bwd(ws.dxn, x_in, st[0], st[1], self.p(pre + "norm1.weight"), self.g(pre + "norm1.weight"),
                              self.g(pre + "norm1.bias"), dres=dx, dx=dx_alt, dx16=ws.dx16 if l > 0 else None,
                              dcolsum=prev_bias, unscale=u, branch_scale=rs(2 * l - 1),
                              branch_drop=self._site(3 * l) if l > 0 else None)
            dx, dx_alt = dx_alt, dx
            self._notify(pre)
        ops.tokens_bwd(dx.view(B, T, D), None if "pos_embed" in self.frozen else self.g("pos_embed"),
                       self.g("cls_token") if d.n_prefix >= 1 else None,
                       self.g("dist_token") if d.n_prefix == 2 else None, ws.dpatch, self.g(self.patch_b),
                       d.n_prefix, unscale=u, drop=self._site(0))
        self._wgrad(ws.dpatch, ws.patches, self.patch_w, B * d.n_patches)
        self._notify("embed")

    def _notify(self, what: str) -> None:
        if self.grad_ready_hook is not None:
            self.grad_ready_hook(what)

    def zero_grad(self) -> None:
        self.flat.grads.zero_()

    def amp_update(self) -> None:
        """Loss-scale bookkeeping when an EXTERNAL optimizer consumes flat.grads (see vitk_amp_update)."""
        if self.dt16 == torch.float16:
            ops.amp_update(self.flat.grads, self.amp, self.amp_scratch, self.growth_interval)
