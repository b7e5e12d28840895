"""Tensor-level wrappers over the C-ABI (raw device pointers + the current CUDA stream).

Every function enqueues kernels of libvitk.so on torch's current stream and returns torch
tensors that merely own the memory.  No function here computes anything with torch ops.
16-bit tensors may be torch.float16 or torch.bfloat16; the wrapper passes the matching vitk_dtype.
"""
from __future__ import annotations

import ctypes as C
import ctypes as C_
import math
from typing import Optional

import torch

from . import _lib
from ._lib import GemmArgs, check

bf16 = torch.bfloat16
f16 = torch.float16
f32 = torch.float32
_DT = {bf16: _lib.DT_BF16, f32: _lib.DT_FP32, f16: _lib.DT_FP16}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _drop(d):
    """(seed int64[1] CUDA tensor, p, site) -> vitk_dropout passed by reference (None: no dropout)."""
    if d is None:
        return None
    seed, p, site = d
    if seed.dtype != torch.int64 or not seed.is_cuda:
        raise TypeError("dropout seed must be a CUDA int64 tensor")
    return C.byref(_lib.Dropout(seed.data_ptr(), float(p), int(site)))


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (libvitk has no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")


def _req16(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (libvitk has no CPU path)")
    if t.dtype not in (bf16, f16):
        raise RuntimeError(f"{name}: expected bfloat16 or float16, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")


def launch_count() -> int:
    return int(_lib.load().vitk_launch_count())


def reset_launch_count() -> None:
    _lib.load().vitk_reset_launch_count()


# --------------------------------------------------------------------------- GEMM
def gemm(A: torch.Tensor, B: torch.Tensor, M: int, N: int, K: int, *, a_mn: bool = False, b_mn: bool = False,
         out: torch.Tensor, bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
         epilogue: int = _lib.EPI_STORE, out2: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None,
         split_k: int = 1, alpha: float = 1.0, alpha_dev: Optional[torch.Tensor] = None, tokens: Optional[tuple] = None,
         pos: Optional[torch.Tensor] = None, lda: Optional[int] = None, ldb: Optional[int] = None,
         colsum_out: Optional[torch.Tensor] = None, row_scale: Optional[torch.Tensor] = None, drop=None) -> torch.Tensor:
    """D[M,N] = A[M,K] @ B[N,K]^T with a fused epilogue (see include/vitk.h).

    A is stored [M,K] (a_mn=False) or [K,M] (a_mn=True); B is stored [N,K] or [K,N].  A and B must share one
    16-bit element type (tcgen05 kind::f16 cannot mix fp16 with bf16)."""
    _req16(A, "gemm A"); _req16(B, "gemm B")
    if A.dtype != B.dtype:
        raise RuntimeError(f"gemm: A ({A.dtype}) and B ({B.dtype}) must have the same 16-bit element type")
    a = GemmArgs()
    a.a_dtype, a.b_dtype = _DT[A.dtype], _DT[B.dtype]
    a.aux_dtype = _DT[aux.dtype] if aux is not None else 0
    a.A, a.B = A.data_ptr(), B.data_ptr()
    a.lda = lda if lda is not None else (M if a_mn else K)
    a.ldb = ldb if ldb is not None else (N if b_mn else K)
    a.a_mn_major, a.b_mn_major = int(a_mn), int(b_mn)
    a.M, a.N, a.K = M, N, K
    a.split_k, a.epilogue = split_k, epilogue
    a.out_dtype = _DT[out.dtype]
    a.alpha = alpha
    a.alpha_dev = _p(alpha_dev)
    if bias is not None:
        _req(bias, f32, "gemm bias")
    if residual is not None:
        _req(residual, f32, "gemm residual")
    a.bias, a.residual, a.ldr = _p(bias), _p(residual), N
    a.out, a.ldo = out.data_ptr(), N
    a.out2, a.ldo2 = _p(out2), N
    a.aux, a.ldaux = _p(aux), N
    if tokens is not None:
        a.rows_per_img, a.tokens_per_img, a.prefix = tokens
        a.pos = _p(pos)
    if colsum_out is not None:
        _req(colsum_out, f32, "gemm colsum_out")
    a.colsum_out = _p(colsum_out)
    if row_scale is not None:
        _req(row_scale, f32, "gemm row_scale")
    a.row_scale = _p(row_scale)
    if drop is not None:       # (seed, p, site): nn.Dropout fused into the epilogue (include/vitk.h)
        a.drop_seed, a.drop_p, a.drop_site = drop[0].data_ptr(), float(drop[1]), int(drop[2])
    check(_lib.load().vitk_gemm(C.byref(a), _stream()), "gemm")
    return out


# --------------------------------------------------------------------------- LayerNorm
def layernorm_fwd(x, gamma, beta, eps: float = 1e-5, y=None, mean=None, rstd=None, dtype=f16):
    _req(x, f32, "layernorm x")
    dim = x.shape[-1]
    rows = x.numel() // dim
    y = torch.empty(x.shape, dtype=dtype, device=x.device) if y is None else y
    mean = torch.empty(rows, dtype=f32, device=x.device) if mean is None else mean
    rstd = torch.empty(rows, dtype=f32, device=x.device) if rstd is None else rstd
    check(_lib.load().vitk_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), _DT[y.dtype],
                                         mean.data_ptr(), rstd.data_ptr(), rows, dim, eps, _stream()), "layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, *, dres=None, dx=None, dx16=None, dcolsum=None, unscale=None,
                  branch_scale=None, branch_drop=None):
    """dx = dres + LN'(dy); dgamma/dbeta/dcolsum += (*unscale) * column sums.  With `branch_scale` ([rows], stochastic depth)
    dx16 and dcolsum carry dx * branch_scale[row]; with `branch_drop` ((seed, p, site)) also the dropout mask of that branch."""
    _req16(dy, "layernorm dy"); _req(x, f32, "layernorm x")
    dim = x.shape[-1]
    rows = x.numel() // dim
    dx = torch.empty_like(x) if dx is None else dx
    check(_lib.load().vitk_layernorm_bwd(dy.data_ptr(), _DT[dy.dtype], x.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                         gamma.data_ptr(), _p(dres), dx.data_ptr(), _p(dx16),
                                         _DT[dx16.dtype] if dx16 is not None else _DT[dy.dtype], dgamma.data_ptr(),
                                         dbeta.data_ptr(), _p(dcolsum), _p(unscale), _p(branch_scale), _drop(branch_drop), rows, dim,
                                         _stream()),
          "layernorm_bwd")
    return dx


# --------------------------------------------------------------------------- attention
def attention_fwd(qkv, B: int, N: int, H: int, scale: float, out=None, lse=None, probs=None, drop=None, q_rows: int = 0):
    """drop = (seed, p, site): training-mode dropout on the attention probabilities (Attention.attn_drop).
    q_rows > 0: only query rows 0..q_rows-1 of out / lse are needed (the other rows may stay unwritten)."""
    _req16(qkv, "attention qkv")
    out = torch.empty(B, N, H * 64, dtype=qkv.dtype, device=qkv.device) if out is None else out
    if out.dtype != qkv.dtype:
        raise RuntimeError("attention: qkv and out must share one element type")
    lse = torch.empty(B, H, N, dtype=f32, device=qkv.device) if lse is None else lse
    if drop is not None:
        if probs is not None:
            raise RuntimeError("attention: probability maps are an eval-mode output; dropout is training-mode only")
        check(_lib.load().vitk_attention_dropout_fwd(qkv.data_ptr(), out.data_ptr(), _DT[qkv.dtype], lse.data_ptr(), B, N, H, scale,
                                                     _drop(drop), _stream()), "attention_dropout_fwd")
        return out, lse
    check(_lib.load().vitk_attention_fwd(qkv.data_ptr(), out.data_ptr(), _DT[qkv.dtype], lse.data_ptr(), _p(probs), B, N, H,
                                         scale, int(q_rows), _stream()), "attention_fwd")
    return out, lse


def attention_bwd(qkv, out, dout, lse, B: int, N: int, H: int, scale: float, dqkv=None, delta=None, drop=None, q_rows: int = 0):
    """q_rows > 0: the caller guarantees dout is zero from query row q_rows on (dqkv is still complete)."""
    _req16(qkv, "attention qkv")
    _req(out, qkv.dtype, "attention out"); _req(dout, qkv.dtype, "attention dout")
    dqkv = torch.empty_like(qkv) if dqkv is None else dqkv
    delta = torch.empty(B, H, N, dtype=f32, device=qkv.device) if delta is None else delta
    if drop is not None:
        check(_lib.load().vitk_attention_dropout_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), delta.data_ptr(),
                                                     dqkv.data_ptr(), _DT[qkv.dtype], B, N, H, scale, _drop(drop), _stream()),
              "attention_dropout_bwd")
        return dqkv
    check(_lib.load().vitk_attention_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), delta.data_ptr(),
                                         dqkv.data_ptr(), _DT[qkv.dtype], B, N, H, scale, int(q_rows), _stream()), "attention_bwd")
    return dqkv


# --------------------------------------------------------------------------- tokens
def patchify(images, P: int, out=None, dtype=f16, channel_last: bool = False):
    """[B,C,H,W] fp32 -> 16-bit patch matrix [B*gh*gw, C*P*P]; k = c*P*P + ky*P + kx (Conv2d weight order), or with
    channel_last k = (ky*P + kx)*C + c (the Rearrange + Linear projection)."""
    _req(images, f32, "patchify images")
    B, Cc, H, W = images.shape
    rows = B * (H // P) * (W // P)
    out = torch.empty(rows, Cc * P * P, dtype=dtype, device=images.device) if out is None else out
    fn = _lib.load().vitk_patchify_hwc if channel_last else _lib.load().vitk_patchify
    check(fn(images.data_ptr(), out.data_ptr(), _DT[out.dtype], B, Cc, H, W, P, _stream()), "patchify")
    return out


def prefix_tokens_fwd(x, cls_tok, dist_tok, pos, n_prefix: int, drop=None):
    B, T, dim = x.shape
    check(_lib.load().vitk_prefix_tokens_fwd(x.data_ptr(), _p(cls_tok), _p(dist_tok), pos.data_ptr(), B, T, dim, n_prefix,
                                             _drop(drop), _stream()), "prefix_tokens_fwd")
    return x


def tokens_bwd(dx, dpos, dcls, ddist, dpatch16, dbias, n_prefix: int, unscale=None, drop=None):
    B, T, dim = dx.shape
    check(_lib.load().vitk_tokens_bwd(dx.data_ptr(), _p(dpos), _p(dcls), _p(ddist), _p(dpatch16),
                                      _DT[dpatch16.dtype] if dpatch16 is not None else 0, _p(dbias), _p(unscale), B, T, dim,
                                      n_prefix, _drop(drop), _stream()), "tokens_bwd")


def gather_rows(src, n: int, out):
    """out[b, j, :] = src[b, j, :] for j < n: the leading n token rows of every image ([B,T,D] -> [B,n,D], any 16/32-bit dtype)."""
    B, T, D = src.shape
    if not (src.is_contiguous() and out.is_contiguous() and out.dtype == src.dtype and out.numel() == B * n * D):
        raise ValueError("gather_rows: contiguous [B,T,D] source and a same-dtype [B,n,D] destination expected")
    check(_lib.load().vitk_gather_rows(src.data_ptr(), out.data_ptr(), B, T, n, D * src.element_size(), _stream()), "gather_rows")
    return out


def expand_rows(src, n: int, out):
    """out[b, j, :] = src[b, j, :] for j < n, zero for the other token rows ([B,n,D] -> dense [B,T,D])."""
    B, T, D = out.shape
    if not (src.is_contiguous() and out.is_contiguous() and out.dtype == src.dtype and src.numel() == B * n * D):
        raise ValueError("expand_rows: contiguous [B,n,D] source and a same-dtype dense [B,T,D] destination expected")
    check(_lib.load().vitk_expand_rows(src.data_ptr(), out.data_ptr(), B, T, n, D * out.element_size(), _stream()), "expand_rows")
    return out


# --------------------------------------------------------------------------- heads
def head_fwd(x, gamma, beta, W0, b0, W1, b1, n_heads: int, eps: float = 1e-5, pooled=None):
    """pooled (optional, fp32 [n_heads,B,dim]) receives norm(x)[:, h] -- the reference's forward_features output."""
    _req(x, f32, "head x")
    B, T, dim = x.shape
    Cc = W0.shape[0]
    logits0 = torch.empty(B, Cc, dtype=f32, device=x.device)
    logits1 = torch.empty(B, Cc, dtype=f32, device=x.device) if n_heads == 2 else None
    xhat = torch.empty(n_heads, B, dim, dtype=f32, device=x.device)
    rstd = torch.empty(n_heads, B, dtype=f32, device=x.device)
    check(_lib.load().vitk_head_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), W0.data_ptr(), _p(b0), _p(W1), _p(b1),
                                    logits0.data_ptr(), _p(logits1), xhat.data_ptr(), rstd.data_ptr(), _p(pooled), B, T, dim, Cc,
                                    n_heads, eps, _stream()), "head_fwd")
    return logits0, logits1, xhat, rstd


def head_bwd(dl0, dl1, xhat, rstd, gamma, beta, W0, W1, dx, dx16, dgamma, dbeta, dW0, db0, dW1, db1, dcolsum,
             T: int, n_heads: int, loss_scale=None, branch_scale=None, branch_drop=None):
    """dx / dx16 = S * dLoss/dx (S = *loss_scale), parameter gradients are true (unscaled)."""
    B, Cc = dl0.shape
    dim = W0.shape[1]
    check(_lib.load().vitk_head_bwd(dl0.data_ptr(), _p(dl1), xhat.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                    W0.data_ptr(), _p(W1), dx.data_ptr(), _p(dx16), _DT[dx16.dtype] if dx16 is not None else 0,
                                    dgamma.data_ptr(), dbeta.data_ptr(), dW0.data_ptr(), _p(db0), _p(dW1), _p(db1), _p(dcolsum),
                                    _p(loss_scale), _p(branch_scale), _drop(branch_drop), B, T, dim, Cc, n_heads, _stream()),
          "head_bwd")


def dropout_mask(seed, p: float, site: int, rows: int, cols: int):
    """fp32 [rows, cols] factors (0 or 1/(1-p)) the fused kernels derive for (seed, site) -- tests / debugging."""
    out = torch.empty(rows, cols, dtype=f32, device=seed.device)
    check(_lib.load().vitk_dropout_mask(_drop((seed, p, site)), out.data_ptr(), rows, cols, _stream()), "dropout_mask")
    return out


def droppath_scale(uniform, drop_prob, T: int, out=None):
    """[branches, B] uniforms + [branches] drop probabilities -> [branches, B*T] per-row stochastic-depth factors."""
    _req(uniform, f32, "droppath uniform"); _req(drop_prob, f32, "droppath drop_prob")
    nb, B = uniform.shape
    out = torch.empty(nb, B * T, dtype=f32, device=uniform.device) if out is None else out
    check(_lib.load().vitk_droppath_scale(uniform.data_ptr(), drop_prob.data_ptr(), out.data_ptr(), nb, B, T, _stream()),
          "droppath_scale")
    return out


# --------------------------------------------------------------------------- loss
def loss_fwd_bwd(cls_logits, dist_logits, teacher_logits, labels, *, mode: int, w_cls: float, w_dist: float, T: float = 1.0,
                 label_smoothing: float = 0.0, grad_div: float = 1.0):
    _req(cls_logits, f32, "loss cls_logits")
    if labels.dtype != torch.int64:
        raise RuntimeError("loss labels must be int64")
    B, Cc = cls_logits.shape
    out = torch.empty(8, dtype=f32, device=cls_logits.device)
    dcls = torch.empty_like(cls_logits)
    ddist = torch.empty_like(dist_logits) if dist_logits is not None else None
    check(_lib.load().vitk_loss_fwd_bwd(cls_logits.data_ptr(), _p(dist_logits), _p(teacher_logits), labels.data_ptr(),
                                        out.data_ptr(), dcls.data_ptr(), _p(ddist), B, Cc, mode, w_cls, w_dist, T,
                                        label_smoothing, grad_div, _stream()), "loss_fwd_bwd")
    return out, dcls, ddist


# --------------------------------------------------------------------------- optimizer / loss scale
_SQ_SCRATCH = {}


def sqnorm_scratch(device) -> torch.Tensor:
    """Per-device partial-sum buffer of the fixed-order gradient norm (zeroed once; the kernel resets its counter itself)."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    buf = _SQ_SCRATCH.get(key)
    if buf is None:
        buf = torch.zeros(int(_lib.load().vitk_sqnorm_scratch_floats()), dtype=f32, device=torch.device("cuda", key))
        _SQ_SCRATCH[key] = buf
    return buf


def grad_sqnorm(grads, state, scratch=None):
    scratch = sqnorm_scratch(grads.device) if scratch is None else scratch
    check(_lib.load().vitk_grad_sqnorm(grads.data_ptr(), grads.numel(), state.data_ptr(), scratch.data_ptr(), _stream()), "grad_sqnorm")


def adamw_step(params, grads, exp_avg, exp_avg_sq, params_bf16, params_fp16, chunk_off, chunk_len, chunk_lr_scale, chunk_wd,
               state, amp_state, beta1: float, beta2: float, eps: float, max_grad_norm: float, growth_interval: int = 2000):
    check(_lib.load().vitk_adamw_step(params.data_ptr(), grads.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                      _p(params_bf16), _p(params_fp16), chunk_off.data_ptr(), chunk_len.data_ptr(),
                                      chunk_lr_scale.data_ptr(), chunk_wd.data_ptr(), chunk_off.numel(), state.data_ptr(),
                                      _p(amp_state), beta1, beta2, eps, max_grad_norm, growth_interval, _stream()), "adamw_step")


def amp_update(grads, amp_state, scratch4, growth_interval: int = 2000):
    check(_lib.load().vitk_amp_update(grads.data_ptr(), grads.numel(), amp_state.data_ptr(), scratch4.data_ptr(),
                                      sqnorm_scratch(grads.device).data_ptr(), int(growth_interval), _stream()), "amp_update")


def cast_bf16(src, dst=None):
    _req(src, f32, "cast src")
    dst = torch.empty(src.shape, dtype=bf16, device=src.device) if dst is None else dst
    check(_lib.load().vitk_cast_f32_to_16(src.data_ptr(), dst.data_ptr(), None, src.numel(), _stream()), "cast")
    return dst


def cast_fp16(src, dst=None):
    _req(src, f32, "cast src")
    dst = torch.empty(src.shape, dtype=f16, device=src.device) if dst is None else dst
    check(_lib.load().vitk_cast_f32_to_16(src.data_ptr(), None, dst.data_ptr(), src.numel(), _stream()), "cast")
    return dst


def colsum16(x, out, unscale=None):
    _req16(x, "colsum x")
    dim = x.shape[-1]
    check(_lib.load().vitk_colsum16(x.data_ptr(), _DT[x.dtype], out.data_ptr(), _p(unscale), x.numel() // dim, dim, _stream()),
          "colsum")
    return out


def ensemble_probs(logits, weights):
    _req(logits, f32, "ensemble logits")
    F, B, Cc = logits.shape
    probs = torch.empty(B, Cc, dtype=f32, device=logits.device)
    pred = torch.empty(B, dtype=torch.int64, device=logits.device)
    check(_lib.load().vitk_ensemble_probs(logits.data_ptr(), weights.data_ptr(), probs.data_ptr(), pred.data_ptr(), F, B, Cc,
                                          _stream()), "ensemble_probs")
    return probs, pred


def attention_rollout(probs, fusion: str = "mean"):
    _req(probs, f32, "rollout probs")
    L, B, H, N, _ = probs.shape
    fus = {"mean": 0, "max": 1, "min": 2}[fusion]
    out = torch.empty(B, N, N, dtype=f32, device=probs.device)
    scratch = torch.empty(2, B, N, N, dtype=f32, device=probs.device)
    check(_lib.load().vitk_attention_rollout(probs.data_ptr(), out.data_ptr(), scratch.data_ptr(), L, B, H, N, fus, _stream()),
          "attention_rollout")
    return out


def attention_rollout_row(probs, row: int = 0, fusion: str = "mean", image_major: bool = False):
    """Row `row` of attention_rollout(probs, fusion) as fp32 [B,N], by L vector-matrix products (maps read once).
    probs: contiguous fp32 [L,B,H,N,N], or [B,L,H,N,N] with image_major=True (one image's maps contiguous)."""
    _req(probs, f32, "rollout probs")
    if image_major:
        B, L, H, N, _ = probs.shape
        sl, sb = H * N * N, L * H * N * N
    else:
        L, B, H, N, _ = probs.shape
        sl, sb = B * H * N * N, H * N * N
    fus = {"mean": 0, "max": 1, "min": 2}[fusion]
    out = torch.empty(B, N, dtype=f32, device=probs.device)
    check(_lib.load().vitk_attention_rollout_row(probs.data_ptr(), out.data_ptr(), sl, sb, L, B, H, N, int(row), fus, _stream()),
          "attention_rollout_row")
    return out


def attention_probs(qkv, lse, B: int, N: int, H: int, scale: float, probs, batch_stride: int = 0):
    """Eval-mode attention maps from the lse of a preceding attention_fwd: image b's [H,N,N] maps are written at
    probs.data_ptr() + b * batch_stride elements (0: contiguous [B,H,N,N])."""
    _req16(qkv, "attention qkv")
    if probs.dtype != f32 or not probs.is_cuda:
        raise TypeError("attention_probs: probs must be a CUDA fp32 tensor")
    stride = int(batch_stride) if batch_stride else H * N * N
    check(_lib.load().vitk_attention_probs(qkv.data_ptr(), _DT[qkv.dtype], lse.data_ptr(), probs.data_ptr(), stride, B, N, H, scale,
                                           _stream()), "attention_probs")
    return probs


# --------------------------------------------------------------------------- on-device metrics
def metrics_update(logits, labels, confusion, scores=None, score_labels=None, count=None):
    """confusion (int64 [C*C+1]) += this batch; optionally appends softmax(logits)[:, 1] / labels at *count (binary AUROC)."""
    _req(logits, f32, "metrics logits")
    B, Cc = logits.shape
    if labels.dtype != torch.int64 or not labels.is_cuda or confusion.dtype != torch.int64 or confusion.numel() != Cc * Cc + 1:
        raise TypeError("metrics_update: labels int64 CUDA [B], confusion int64 [C*C+1]")
    cap = scores.numel() if scores is not None else 0
    check(_lib.load().vitk_metrics_update(logits.data_ptr(), labels.data_ptr(), B, Cc, confusion.data_ptr(), _p(scores),
                                          _p(score_labels), _p(count), cap, _stream()), "metrics_update")


def binary_auroc(scores, score_labels, count, scratch=None, out=None):
    """float64 [4] = {auroc, P, N, ties} over the first min(*count, capacity) appended samples."""
    scratch = torch.empty(4, dtype=torch.int64, device=scores.device) if scratch is None else scratch
    out = torch.empty(4, dtype=torch.float64, device=scores.device) if out is None else out
    check(_lib.load().vitk_binary_auroc(scores.data_ptr(), score_labels.data_ptr(), count.data_ptr(), scores.numel(),
                                        scratch.data_ptr(), out.data_ptr(), _stream()), "binary_auroc")
    return out


# --------------------------------------------------------------------------- GPU-side input pipeline
def resize_u16(raw, H: int, W: int, out=None):
    """raw uint16 [B, Hs, Ws] (CUDA) -> gray fp32 [B, H, W] in [0, 1] (cv2.resize INTER_LINEAR semantics, / 65535)."""
    if raw.dtype != torch.uint16 or not raw.is_cuda or raw.dim() != 3 or not raw.is_contiguous():
        raise TypeError("resize_u16: raw must be a contiguous CUDA uint16 tensor [B, Hs, Ws]")
    B, Hs, Ws = raw.shape
    out = torch.empty(B, H, W, dtype=f32, device=raw.device) if out is None else out
    check(_lib.load().vitk_resize_u16(raw.data_ptr(), out.data_ptr(), B, Hs, Ws, H, W, _stream()), "resize_u16")
    return out


def percentile_bounds(x, q_lo: float, q_hi: float, out=None):
    """x fp32 [B, ...] -> fp32 [B, 2] = per-image (torch.quantile(q_lo), torch.quantile(q_hi))."""
    _req(x, f32, "percentile x")
    B = x.shape[0]
    n = x.numel() // B
    out = torch.empty(B, 2, dtype=f32, device=x.device) if out is None else out
    check(_lib.load().vitk_percentile_bounds(x.data_ptr(), B, n, float(q_lo), float(q_hi), out.data_ptr(), _stream()),
          "percentile_bounds")
    return out


def finish_tiles(gray, C: int, *, bounds=None, mean=None, std=None, perm=None, cutmix: bool = False, lam: float = 1.0,
                 box=(0, 0, 0, 0), out=None):
    """gray fp32 [B, H, W] -> fp32 [B, C, H, W]: optional percentile clamp-normalise, channel replicate + Normalize, MixUp/CutMix.
    box = (x1, y1, x2, y2) in the reference's CutMix naming (columns x, rows y)."""
    _req(gray, f32, "finish gray")
    B, H, W = gray.shape
    out = torch.empty(B, C, H, W, dtype=f32, device=gray.device) if out is None else out
    if (mean is None) != (std is None):
        raise ValueError("finish_tiles: mean and std come together")
    marr = sarr = None
    if mean is not None:
        if len(mean) != C or len(std) != C:
            raise ValueError("finish_tiles: one mean / std per output channel")
        import ctypes
        marr = (ctypes.c_float * C)(*[float(v) for v in mean])
        sarr = (ctypes.c_float * C)(*[float(v) for v in std])
    if perm is not None and (perm.dtype != torch.int32 or not perm.is_cuda or perm.numel() != B):
        raise TypeError("finish_tiles: perm must be a CUDA int32 tensor [B]")
    x1, y1, x2, y2 = (int(v) for v in box)
    check(_lib.load().vitk_finish_tiles(gray.data_ptr(), _p(bounds), out.data_ptr(), B, C, H, W, marr, sarr, _p(perm),
                                        int(bool(cutmix)), float(lam), x1, y1, x2, y2, _stream()), "finish_tiles")
    return out


_TILE_KIND = {f32: 0, torch.uint16: 1, f16: 2, bf16: 3}


def tiles_to_patches(tiles, C: int, P: int, *, bounds=None, mean=None, std=None, out=None, dtype=f16):
    """single-channel tiles [B,H,W] or [B,1,H,W] (fp32 in [0,1] | raw uint16 | fp16 | bf16, CUDA) -> 16-bit patch matrix
    [B*(H/P)*(W/P), C*P*P] (channel-first patch vectors): vitk_finish_tiles + vitk_patchify in one pass."""
    if tiles.dim() == 4 and tiles.shape[1] == 1:
        tiles = tiles[:, 0]
    if tiles.dim() != 3 or not tiles.is_cuda or not tiles.is_contiguous() or tiles.dtype not in _TILE_KIND:
        raise TypeError("tiles_to_patches: tiles must be a contiguous CUDA [B,H,W] tensor of fp32 / uint16 / fp16 / bf16")
    B, H, W = tiles.shape
    if (mean is None) != (std is None):
        raise ValueError("tiles_to_patches: mean and std come together")
    marr = sarr = None
    if mean is not None:
        if len(mean) != C or len(std) != C:
            raise ValueError("tiles_to_patches: one mean / std per output channel")
        marr = (C_.c_float * C)(*[float(v) for v in mean])
        sarr = (C_.c_float * C)(*[float(v) for v in std])
    if bounds is not None:
        _req(bounds, f32, "tiles_to_patches bounds")
    rows, kdim = B * (H // P) * (W // P), C * P * P
    out = torch.empty(rows, kdim, dtype=dtype, device=tiles.device) if out is None else out
    _req16(out, "tiles_to_patches out")
    if out.numel() != rows * kdim:
        raise ValueError("tiles_to_patches: out has the wrong size")
    check(_lib.load().vitk_tiles_to_patches(tiles.data_ptr(), _TILE_KIND[tiles.dtype], _p(bounds), marr, sarr, out.data_ptr(),
                                            _DT[out.dtype], B, C, H, W, P, _stream()), "tiles_to_patches")
    return out


# --------------------------------------------------------------------------- general classification tail (gap pooling, pre_logits)
def pool_norm_fwd(x, gamma, beta, t0: int, t1: int, eps: float = 1e-5):
    """x fp32 [B,T,D] -> (pooled [B,D] = mean over tokens [t0,t1) of LayerNorm(x), mean [B,T], rstd [B,T])."""
    _req(x, f32, "pool x")
    B, T, dim = x.shape
    pooled = torch.empty(B, dim, dtype=f32, device=x.device)
    mean = torch.empty(B, T, dtype=f32, device=x.device)
    rstd = torch.empty(B, T, dtype=f32, device=x.device)
    check(_lib.load().vitk_pool_norm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), pooled.data_ptr(), mean.data_ptr(),
                                         rstd.data_ptr(), B, T, dim, t0, t1, eps, _stream()), "pool_norm_fwd")
    return pooled, mean, rstd


def pool_norm_bwd(dpooled, x, mean, rstd, gamma, dx, dx16, dgamma, dbeta, dcolsum, t0: int, t1: int, loss_scale=None,
                  branch_scale=None, branch_drop=None):
    B, T, dim = x.shape
    check(_lib.load().vitk_pool_norm_bwd(dpooled.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                         dx.data_ptr(), _p(dx16), _DT[dx16.dtype] if dx16 is not None else 0, dgamma.data_ptr(),
                                         dbeta.data_ptr(), _p(dcolsum), _p(loss_scale), _p(branch_scale), _drop(branch_drop),
                                         B, T, dim, t0, t1, _stream()), "pool_norm_bwd")


def dense_fwd(x, W, bias, act: int = 0):
    _req(x, f32, "dense x")
    B, in_dim = x.shape
    out_dim = W.shape[0]
    y = torch.empty(B, out_dim, dtype=f32, device=x.device)
    check(_lib.load().vitk_dense_fwd(x.data_ptr(), W.data_ptr(), _p(bias), y.data_ptr(), B, in_dim, out_dim, act, _stream()),
          "dense_fwd")
    return y


def dense_bwd(dy, y, x, W, dW, db, act: int = 0, need_dx: bool = True):
    """Returns dx [B, in_dim] (or None); dW / db are accumulated."""
    B, in_dim = x.shape
    out_dim = W.shape[0]
    dz = torch.empty(B, out_dim, dtype=f32, device=x.device)
    dx = torch.empty(B, in_dim, dtype=f32, device=x.device) if need_dx else None
    check(_lib.load().vitk_dense_bwd(dy.data_ptr(), _p(y), x.data_ptr(), W.data_ptr(), dz.data_ptr(), _p(dx), dW.data_ptr(),
                                     _p(db), B, in_dim, out_dim, act, _stream()), "dense_bwd")
    return dx


# --------------------------------------------------------------------------- frozen-teacher fast path
def affine_relu_nhwc(x, C: int, scale, shift, out=None, relu: bool = True, out_channel_offset: int = 0):
    """x: contiguous 16-bit NHWC tensor [..., C_total]; its first C channels go through y = max(0, x * scale + shift) into
    channels [out_channel_offset, out_channel_offset + C) of `out` (contiguous [..., C_out_total]; default: a new compact [..., C])."""
    _req16(x, "affine_relu x")
    pixels = x.numel() // x.shape[-1]
    if out is None:
        out = torch.empty(*x.shape[:-1], C, dtype=x.dtype, device=x.device)
    _req(out, x.dtype, "affine_relu out")
    if out.numel() // out.shape[-1] != pixels:
        raise RuntimeError("affine_relu_nhwc: out must have the pixel count of x")
    off = int(out_channel_offset)
    if off < 0 or off % 8 or off + C > out.shape[-1]:
        raise ValueError("affine_relu_nhwc: out_channel_offset must be a multiple of 8 with offset + C <= channels of out")
    _req(scale, f32, "affine_relu scale"); _req(shift, f32, "affine_relu shift")
    check(_lib.load().vitk_affine_relu_nhwc(x.data_ptr(), x.shape[-1], out.data_ptr() + off * out.element_size(), out.shape[-1],
                                            scale.data_ptr(), shift.data_ptr(), pixels, C, _DT[x.dtype], int(relu), _stream()),
          "affine_relu_nhwc")
    return out


def im2col_rows(x_nhwc, kernel: int, stride: int, pad: int, out=None):
    """16-bit NHWC [B,H,W,C] -> patch rows [B*OH*OW, ld] (ld = k*k*C rounded up to 8), element order (ky, kx, c)."""
    _req16(x_nhwc, "im2col x")
    B, H, W, Cc = x_nhwc.shape
    OH, OW = (H + 2 * pad - kernel) // stride + 1, (W + 2 * pad - kernel) // stride + 1
    ld = (kernel * kernel * Cc + 7) // 8 * 8
    out = torch.empty(B * OH * OW, ld, dtype=x_nhwc.dtype, device=x_nhwc.device) if out is None else out
    check(_lib.load().vitk_im2col_rows(x_nhwc.data_ptr(), out.data_ptr(), B, H, W, Cc, kernel, stride, pad, ld, _stream()), "im2col_rows")
    return out


def stem_conv7_weights(weight, dtype):
    """Conv2d weight [64, 3, 7, 7] (norm0 already folded in) -> the [64, 192] filter matrix of vitk_stem_conv7:
    column ky*24 + 1 + kx*3 + c, zeros in slots 0, 22, 23 of every ky segment and from column 168 on."""
    if tuple(weight.shape) != (64, 3, 7, 7):
        raise ValueError("stem_conv7_weights: a [64, 3, 7, 7] filter bank expected")
    w = weight.detach().float().permute(0, 2, 3, 1).reshape(64, 7, 21)
    w = torch.nn.functional.pad(w, (1, 2)).reshape(64, 168)
    return torch.nn.functional.pad(w, (0, 24)).to(dtype).contiguous()


def stem_conv7(x_nhwc, w, bias, relu: bool = True, out=None):
    """7x7 / stride 2 / padding 3 convolution of a 16-bit NHWC [B,H,W,3] batch (implicit GEMM, no patch matrix):
    -> [B,H/2,W/2,64]; w from stem_conv7_weights, bias fp32 [64]."""
    _req16(x_nhwc, "stem x")
    B, H, W, Cc = x_nhwc.shape
    if Cc != 3 or not x_nhwc.is_contiguous():
        raise ValueError("stem_conv7: a contiguous NHWC batch with 3 channels expected")
    if w.dtype != x_nhwc.dtype or not w.is_cuda or not w.is_contiguous() or tuple(w.shape) != (64, 192):
        raise TypeError("stem_conv7: w must be a contiguous CUDA [64, 192] tensor of x's dtype (stem_conv7_weights)")
    _req(bias, f32, "stem bias")
    out = torch.empty(B, H // 2, W // 2, 64, dtype=x_nhwc.dtype, device=x_nhwc.device) if out is None else out
    check(_lib.load().vitk_stem_conv7(x_nhwc.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), B, H, W, int(bool(relu)),
                                      _DT[x_nhwc.dtype], _stream()), "stem_conv7")
    return out


def dense_bottleneck(x, C: int, scale, shift, w, bias, out=None):
    """x: 16-bit NHWC [..., Ct] (the first C channels are read), w: 16-bit [128, C], scale / shift fp32 [C], bias fp32 [128]
    -> relu(relu(x[..., :C] * scale + shift) @ w^T + bias) as 16-bit [..., 128]."""
    _req16(x, "bottleneck x")
    ct = x.shape[-1]
    pixels = x.numel() // ct
    if w.dtype != x.dtype or not w.is_cuda or not w.is_contiguous() or tuple(w.shape) != (128, C):
        raise TypeError("dense_bottleneck: w must be a contiguous CUDA [128, C] tensor of x's dtype")
    _req(scale, f32, "bottleneck scale"); _req(shift, f32, "bottleneck shift"); _req(bias, f32, "bottleneck bias")
    if scale.numel() != C or shift.numel() != C or bias.numel() != 128:
        raise ValueError("dense_bottleneck: scale / shift need C entries, bias 128")
    out = torch.empty(*x.shape[:-1], 128, dtype=x.dtype, device=x.device) if out is None else out
    check(_lib.load().vitk_dense_bottleneck(x.data_ptr(), ct, scale.data_ptr(), shift.data_ptr(), w.data_ptr(), bias.data_ptr(),
                                            out.data_ptr(), pixels, C, _DT[x.dtype], _stream()), "dense_bottleneck")
    return out


def pool_nhwc(x, out, kernel: int, stride: int, pad: int, is_max: bool):
    """x: contiguous 16-bit [B,H,W,C]; out: contiguous [B,OH,OW,C_total >= C] whose first C channels receive the pooled map."""
    _req16(x, "pool x"); _req(out, x.dtype, "pool out")
    B, H, W, Cc = x.shape
    OH, OW = (H + 2 * pad - kernel) // stride + 1, (W + 2 * pad - kernel) // stride + 1
    if tuple(out.shape[:3]) != (B, OH, OW) or out.shape[3] < Cc:
        raise RuntimeError(f"pool_nhwc: out must be [B={B},{OH},{OW},>= {Cc}], got {tuple(out.shape)}")
    check(_lib.load().vitk_pool_nhwc(x.data_ptr(), out.data_ptr(), out.shape[3], B, H, W, Cc, kernel, stride, pad, int(is_max),
                                     _DT[x.dtype], _stream()), "pool_nhwc")
    return out


def cls_attention_heatmap(src, out_hw, n_prefix: int = 1):
    """Class-token heat maps fp32 [B,out_h,out_w] (attention_utils.py:50-67, every image of the batch).
    src: one layer's maps fp32 [B,H,N,N] (any batch / head strides, rows contiguous), a rollout row [B,N] (then n_prefix
    patch-prefix columns are skipped) or a grid [B,g,g] (n_prefix is ignored)."""
    if not src.is_cuda:
        raise RuntimeError("heat-map source: expected a CUDA tensor (libvitk has no CPU path)")
    if src.dtype != f32:
        raise RuntimeError(f"heat-map source: expected dtype {f32}, got {src.dtype}")
    oh, ow = int(out_hw[0]), int(out_hw[1])
    if src.dim() == 4:
        B, H, N, N2 = src.shape
        if N != N2 or src.stride(3) != 1:
            raise ValueError("attention maps must be [B,H,N,N] with contiguous rows")
        sb, sh, cols = src.stride(0), src.stride(1), N - n_prefix
    elif src.dim() == 2:
        if src.stride(1) != 1:
            raise ValueError("rollout rows must be contiguous")
        B, H, sb, sh, cols = src.shape[0], 1, src.stride(0), 0, src.shape[1] - n_prefix
    elif src.dim() == 3:
        src = src.contiguous()
        B, H, sb, sh, cols, n_prefix = src.shape[0], 1, src.shape[1] * src.shape[2], 0, src.shape[1] * src.shape[2], 0
    else:
        raise ValueError("heat-map source must be [B,H,N,N], [B,N] or [B,g,g]")
    g = math.isqrt(max(cols, 0))
    if g == 0 or g * g != cols:
        raise ValueError(f"{cols} patch columns are not a square grid (n_prefix={n_prefix}; a distilled DeiT has 2 prefix tokens)")
    out = torch.empty(B, oh, ow, dtype=f32, device=src.device)
    check(_lib.load().vitk_cls_attention_heatmap(src.data_ptr(), out.data_ptr(), sb, sh, B, H, int(n_prefix), g, oh, ow, _stream()),
          "cls_attention_heatmap")
    return out
