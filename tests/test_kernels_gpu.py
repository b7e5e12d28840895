"""Kernel-level parity: every C-ABI entry point of libvitk.so against a plain PyTorch fp32
reference of the same op, on the GPU.  Tolerances are stated per test (16-bit operands, fp32
accumulation; fp16 has an 11-bit significand, bf16 an 8-bit one)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from thyroid_vit_cnn_comparison_b200 import _lib, ops  # noqa: E402

DEV = "cuda"
F16, BF16 = torch.float16, torch.bfloat16
OUT_TOL = {F16: 7e-4, BF16: 6e-3}       # relative L2 of a 16-bit rounded output


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _rand(*shape, scale=1.0, dtype=torch.float32, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


# ------------------------------------------------------------------ GEMM (tcgen05)
GEMM_SHAPES = [
    (256, 192, 192), (384, 576, 192), (6336, 576, 192), (6336, 192, 768), (6336, 768, 192),
    (198, 192, 192), (1000, 64, 64), (777, 2304, 768), (512, 256, 3072), (130, 136, 72),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("dt", [F16, BF16])
def test_gemm_tn_bias(M, N, K, dt):
    A = _rand(M, K, dtype=dt, seed=1)
    W = _rand(N, K, scale=0.05, dtype=dt, seed=2)
    bias = _rand(N, seed=3)
    out = torch.empty(M, N, dtype=dt, device=DEV)
    ops.gemm(A, W, M, N, K, out=out, bias=bias)
    ref = A.float() @ W.float().t() + bias
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < OUT_TOL[dt]


def test_gemm_rejects_mixed_operand_formats():
    """tcgen05 kind::f16 with a_format != b_format was probed on B200: cudaErrorIllegalInstruction.  The C-ABI
    therefore refuses the combination up front instead of faulting the context."""
    A, W = _rand(128, 64, dtype=F16), _rand(64, 64, dtype=BF16)
    with pytest.raises(RuntimeError):
        ops.gemm(A, W, 128, 64, 64, out=torch.empty(128, 64, device=DEV))


@pytest.mark.parametrize("M,N,K", [(6336, 192, 192), (1000, 192, 768), (130, 136, 72)])
def test_gemm_residual_fp32(M, N, K):
    A = _rand(M, K, dtype=F16, seed=1)
    W = _rand(N, K, scale=0.05, dtype=F16, seed=2)
    bias = _rand(N, seed=3)
    res = _rand(M, N, seed=4)
    out = torch.empty(M, N, dtype=torch.float32, device=DEV)
    ops.gemm(A, W, M, N, K, out=out, bias=bias, residual=res)
    ref = A.float() @ W.float().t() + bias + res
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 1e-5
    assert (out - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("M,N,K", [(6336, 768, 192), (300, 256, 128)])
@pytest.mark.parametrize("dt", [F16, BF16])
def test_gemm_gelu(M, N, K, dt):
    A = _rand(M, K, dtype=dt, seed=1)
    W = _rand(N, K, scale=0.1, dtype=dt, seed=2)
    bias = _rand(N, seed=3)
    dact = torch.empty(M, N, dtype=dt, device=DEV)
    act = torch.empty(M, N, dtype=dt, device=DEV)
    ops.gemm(A, W, M, N, K, out=dact, out2=act, bias=bias, epilogue=_lib.EPI_GELU)
    ref = (A.float() @ W.float().t() + bias).requires_grad_(True)
    ref_act = torch.nn.functional.gelu(ref)          # exact erf, as nn.GELU() (vision_transformer_base.py:212-219)
    ref_act.sum().backward()                         # d gelu / d pre
    torch.cuda.synchronize()
    assert rel_l2(act, ref_act) < OUT_TOL[dt]
    assert rel_l2(dact, ref.grad) < OUT_TOL[dt]
    assert (act.float() - ref_act).abs().max().item() < {F16: 4e-3, BF16: 6e-2}[dt]


@pytest.mark.parametrize("M,N,K", [(6336, 768, 192), (6336, 192, 576), (333, 3072, 768), (130, 72, 136)])
def test_gemm_dgrad_b_mn_major(M, N, K):
    """dX[M,N] = dY[M,K] @ W[K,N]  -- W read MN-major exactly as stored by nn.Linear ([out=K, in=N])."""
    dY = _rand(M, K, dtype=F16, seed=1)
    W = _rand(K, N, scale=0.05, dtype=F16, seed=2)
    out = torch.empty(M, N, dtype=F16, device=DEV)
    ops.gemm(dY, W, M, N, K, b_mn=True, out=out)
    ref = dY.float() @ W.float()
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < OUT_TOL[F16]


@pytest.mark.parametrize("dt", [F16, BF16])
def test_gemm_dgelu(dt):
    M, N, K = 1000, 768, 192
    dY = _rand(M, K, dtype=dt, seed=1)
    W = _rand(K, N, scale=0.05, dtype=dt, seed=2)
    dact = _rand(M, N, dtype=dt, seed=5)            # the derivative saved by the forward GELU epilogue
    out = torch.empty(M, N, dtype=dt, device=DEV)
    ops.gemm(dY, W, M, N, K, b_mn=True, out=out, aux=dact, epilogue=_lib.EPI_DGELU)
    ref = (dY.float() @ W.float()) * dact.float()
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < OUT_TOL[dt]


@pytest.mark.parametrize("Mc,No,Ko,split", [(6336, 576, 192, 8), (6336, 192, 768, 16), (1000, 768, 192, 3),
                                            (50688, 576, 192, 29), (198, 136, 72, 1)])
def test_gemm_wgrad_mn_mn_splitk(Mc, No, Ko, split):
    """dW[No,Ko] += u * dY[Mc,No]^T @ X[Mc,Ko]: both operands MN-major, split-K with fp32 atomics, device-side 1/S."""
    dY = _rand(Mc, No, dtype=F16, seed=1)
    X = _rand(Mc, Ko, dtype=F16, seed=2)
    out = torch.zeros(No, Ko, dtype=torch.float32, device=DEV)
    u = torch.tensor([0.25], device=DEV)
    ops.gemm(dY, X, No, Ko, Mc, a_mn=True, b_mn=True, out=out, split_k=split, epilogue=_lib.EPI_ATOMIC_ADD, alpha_dev=u)
    ref = 0.25 * (dY.float().t() @ X.float())
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 1e-4


@pytest.mark.parametrize("Mc,No,Ko,split", [(6336, 576, 192, 8), (6336, 768, 192, 5), (50688, 2304, 768, 2), (1000, 136, 72, 3)])
@pytest.mark.parametrize("dt", [F16, BF16])
def test_gemm_wgrad_with_bias_grad(Mc, No, Ko, split, dt):
    """The bias gradient (column sums of dY) falls out of the wgrad GEMM: one extra N=16 MMA per k-step against ones."""
    dY = _rand(Mc, No, dtype=dt, seed=1)
    X = _rand(Mc, Ko, dtype=dt, seed=2)
    out = torch.zeros(No, Ko, dtype=torch.float32, device=DEV)
    db = torch.full((No,), 0.5, dtype=torch.float32, device=DEV)      # accumulates (+=)
    u = torch.tensor([0.25], device=DEV)
    ops.gemm(dY, X, No, Ko, Mc, a_mn=True, b_mn=True, out=out, split_k=split, epilogue=_lib.EPI_ATOMIC_ADD, alpha_dev=u,
             colsum_out=db)
    torch.cuda.synchronize()
    assert rel_l2(out, 0.25 * (dY.float().t() @ X.float())) < {F16: 1e-4, BF16: 1e-4}[dt]
    assert rel_l2(db, 0.5 + 0.25 * dY.float().sum(0)) < 1e-4


@pytest.mark.parametrize("M,N,K", [(6336, 192, 192), (1000, 768, 3072)])
def test_gemm_residual_row_scale(M, N, K):
    """out = residual + row_scale[m] * (A W^T + bias): stochastic depth in the residual epilogue (DropPath, :56-64, :283-284)."""
    A = _rand(M, K, dtype=F16, seed=1)
    W = _rand(N, K, scale=0.05, dtype=F16, seed=2)
    bias, res = _rand(N, seed=3), _rand(M, N, seed=4)
    g = torch.Generator(device="cpu").manual_seed(5)
    rs = (torch.floor(0.8 + torch.rand(M, generator=g)) / 0.8).to(DEV)
    out = torch.empty(M, N, device=DEV)
    ops.gemm(A, W, M, N, K, out=out, bias=bias, residual=res, row_scale=rs)
    ref = res + rs[:, None] * (A.float() @ W.float().t() + bias)
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 1e-5
    assert torch.equal(out[rs == 0], res[rs == 0])          # dropped samples keep the residual bit for bit


def test_droppath_scale_kernel():
    nb, B, T = 6, 64, 198
    u = torch.rand(nb, B, device=DEV)
    p = torch.tensor([0.0, 0.02, 0.1, 0.25, 0.5, 0.9], device=DEV)
    s = ops.droppath_scale(u, p, T)
    ref = (torch.floor((1 - p)[:, None] + u) / (1 - p)[:, None])[:, :, None].expand(nb, B, T).reshape(nb, B * T)
    torch.cuda.synchronize()
    assert torch.equal(s, ref)


def test_gemm_tokens_epilogue():
    B, P, T, prefix, D, K = 5, 196, 198, 2, 192, 768
    A = _rand(B * P, K, dtype=F16, seed=1)
    W = _rand(D, K, scale=0.05, dtype=F16, seed=2)
    bias = _rand(D, seed=3)
    pos = _rand(T, D, seed=4)
    x = torch.zeros(B, T, D, dtype=torch.float32, device=DEV)
    ops.gemm(A, W, B * P, D, K, out=x, bias=bias, epilogue=_lib.EPI_TOKENS, tokens=(P, T, prefix), pos=pos)
    ref = torch.zeros_like(x)
    ref[:, prefix:] = (A.float() @ W.float().t() + bias).view(B, P, D) + pos[prefix:]
    torch.cuda.synchronize()
    assert (x - ref).abs().max().item() < 1e-3


# ------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("rows,dim", [(6336, 192), (197 * 3, 768), (50, 384), (33, 1024), (7, 64)])
@pytest.mark.parametrize("dt", [F16, BF16])
def test_layernorm_fwd_bwd(rows, dim, dt):
    x = _rand(rows, dim, seed=1) * 2 + 0.5
    g = _rand(dim, seed=2) * 0.2 + 1
    b = _rand(dim, seed=3) * 0.1
    y, mean, rstd = ops.layernorm_fwd(x, g, b, dtype=dt)
    ref = torch.nn.functional.layer_norm(x, (dim,), g, b, 1e-5)
    assert y.dtype == dt and rel_l2(y, ref) < OUT_TOL[dt]
    S = 8.0                                       # loss scale carried by dy / dres / dx; parameter grads come out unscaled
    dy_true = _rand(rows, dim, seed=4)
    dres_true = _rand(rows, dim, seed=5)
    dy = (dy_true * S).to(dt)
    dg = torch.zeros(dim, device=DEV); db = torch.zeros(dim, device=DEV); dc = torch.zeros(dim, device=DEV)
    dx16 = torch.empty(rows, dim, dtype=dt, device=DEV)
    u = torch.tensor([1.0 / S], device=DEV)
    dx = ops.layernorm_bwd(dy, x, mean, rstd, g, dg, db, dres=dres_true * S, dx16=dx16, dcolsum=dc, unscale=u)
    xr = x.clone().requires_grad_(True); gr = g.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (dim,), gr, br, 1e-5).backward(dy.float() / S)
    ref_dx = xr.grad + dres_true
    torch.cuda.synchronize()
    assert rel_l2(dx / S, ref_dx) < 1e-5
    assert rel_l2(dx16.float() / S, ref_dx) < OUT_TOL[dt]
    assert rel_l2(dg, gr.grad) < 1e-4
    assert rel_l2(db, br.grad) < 1e-4
    assert rel_l2(dc, ref_dx.sum(0)) < 1e-4
    # stochastic depth: the 16-bit copy and the column sum carry dx * branch_scale[row]; dx itself is unscaled
    gen = torch.Generator(device="cpu").manual_seed(9)
    bs = (torch.floor(0.7 + torch.rand(rows, generator=gen)) / 0.7).to(DEV)
    dg.zero_(); db.zero_(); dc.zero_()
    dx2 = ops.layernorm_bwd(dy, x, mean, rstd, g, dg, db, dres=dres_true * S, dx16=dx16, dcolsum=dc, unscale=u, branch_scale=bs)
    torch.cuda.synchronize()
    assert torch.equal(dx2, dx)
    assert rel_l2(dx16.float() / S, ref_dx * bs[:, None]) < OUT_TOL[dt]
    assert rel_l2(dc, (ref_dx * bs[:, None]).sum(0)) < 1e-4


# ------------------------------------------------------------------ attention
def _attn_ref(qkv, B, N, H, scale):
    q, k, v = qkv.float().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).unbind(0)
    s = (q @ k.transpose(-2, -1)) * scale
    p = s.softmax(-1)
    o = (p @ v).transpose(1, 2).reshape(B, N, H * 64)
    return o, p, torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,N,H", [(2, 198, 3), (3, 197, 12), (1, 64, 1), (2, 577, 3), (1, 17, 2), (2, 257, 2), (1, 785, 2),
                                   (2, 1025, 1), (1, 2305, 1)])
@pytest.mark.parametrize("dt", [F16, BF16])
def test_attention_fwd_bwd(B, N, H, dt):
    scale = 64 ** -0.5
    tol = {F16: 1e-3, BF16: 8e-3}[dt]
    qkv = _rand(B, N, 3 * H * 64, dtype=dt, seed=1)
    out, lse = ops.attention_fwd(qkv, B, N, H, scale)
    qr = qkv.float().requires_grad_(True)
    o_ref, p_ref, lse_ref = _attn_ref(qr, B, N, H, scale)
    torch.cuda.synchronize()
    assert out.dtype == dt and rel_l2(out, o_ref) < tol
    assert (lse - lse_ref).abs().max().item() < 2e-3
    dout = _rand(B, N, H * 64, dtype=dt, seed=2)
    dqkv = ops.attention_bwd(qkv, out, dout, lse, B, N, H, scale)
    o_ref.backward(dout.float())
    torch.cuda.synchronize()
    assert rel_l2(dqkv, qr.grad) < 2 * tol


@pytest.mark.parametrize("B,N,H", [(2, 198, 3), (1, 64, 1), (2, 577, 3), (1, 17, 2), (1, 1025, 2)])
@pytest.mark.parametrize("dt", [F16, BF16])
def test_attention_dropout_fwd_bwd(B, N, H, dt):
    """vitk_attention_dropout_{fwd,bwd}: nn.Dropout on the softmax output (vision_transformer_base.py:184).  The mask is the
    library's counter-based one, exported with vitk_dropout_mask and replayed into an fp32 torch restatement."""
    scale, p, site = 64 ** -0.5, 0.25, 1003
    tol = {F16: 1e-3, BF16: 8e-3}[dt]
    qkv = _rand(B, N, 3 * H * 64, dtype=dt, seed=1)
    seed = torch.tensor([12345], dtype=torch.int64, device=DEV)
    npad = (N + 7) // 8 * 8
    mask = ops.dropout_mask(seed, p, site, B * H * N, npad).view(B, H, N, npad)[..., :N]
    if mask.numel() > 50000:                              # drop frequency (the 17-token case is too small a sample)
        assert abs((mask == 0).float().mean().item() - p) < 0.02
    assert all(v == 0.0 or abs(v - 1.0 / (1.0 - p)) < 1e-6 for v in mask.unique().tolist())
    out, lse = ops.attention_fwd(qkv, B, N, H, scale, drop=(seed, p, site))
    qr = qkv.float().requires_grad_(True)
    q, k, v = qr.view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).unbind(0)
    sc = (q @ k.transpose(-2, -1)) * scale
    o_ref = ((sc.softmax(-1) * mask) @ v).transpose(1, 2).reshape(B, N, H * 64)
    torch.cuda.synchronize()
    assert rel_l2(out, o_ref) < tol
    assert (lse - torch.logsumexp(sc, -1)).abs().max().item() < 2e-3           # lse is that of the unmasked probabilities
    dout = _rand(B, N, H * 64, dtype=dt, seed=2)
    dqkv = ops.attention_bwd(qkv, out, dout, lse, B, N, H, scale, drop=(seed, p, site))
    o_ref.backward(dout.float())
    torch.cuda.synchronize()
    assert rel_l2(dqkv, qr.grad) < 2 * tol
    out2, _ = ops.attention_fwd(qkv, B, N, H, scale, drop=(seed + 1, p, site))  # another seed: another mask
    assert not torch.equal(out2, out)
    with pytest.raises(RuntimeError):
        ops.attention_fwd(qkv, B, N, H, scale, drop=(seed, 1.0, site))


def test_attention_probs_with_batch_stride_equal_the_contiguous_maps():
    B, N, H = 5, 198, 3
    qkv = _rand(B, N, 3 * H * 64, seed=8).to(torch.float16)
    ref = torch.empty(B, H, N, N, device=DEV)
    out, lse = ops.attention_fwd(qkv, B, N, H, 0.125, probs=ref)
    L = 4
    im = torch.zeros(B, L, H, N, N, device=DEV)
    ops.attention_probs(qkv, lse, B, N, H, 0.125, im[:, 2], batch_stride=L * H * N * N)
    torch.cuda.synchronize()
    assert torch.equal(im[:, 2], ref) and im[:, 1].abs().max().item() == 0 and im[:, 3].abs().max().item() == 0


@pytest.mark.parametrize("B,N,H", [(2, 198, 3), (2, 257, 2), (2, 577, 3), (1, 785, 2), (1, 1025, 1)])
def test_attention_probs_rows_sum_to_one(B, N, H):
    """Eval-mode maps on the tensor core for every sequence length (one key tile up to 256 tokens, key tiles beyond)."""
    qkv = _rand(B, N, 3 * H * 64, dtype=F16, seed=1)
    probs = torch.empty(B, H, N, N, device=DEV)
    ops.attention_fwd(qkv, B, N, H, 0.125, probs=probs)
    _, p_ref, _ = _attn_ref(qkv, B, N, H, 0.125)
    torch.cuda.synchronize()
    assert (probs.sum(-1) - 1).abs().max().item() < 1e-5      # reference test_attention_quality.py:96-102
    assert probs.min().item() >= 0 and probs.max().item() <= 1
    assert (probs - p_ref).abs().max().item() < 1e-5


@pytest.mark.parametrize("B,N,H", [(2, 198, 3), (2, 577, 2), (1, 785, 1), (1, 1025, 1)])
def test_attention_outputs_stay_inside_their_buffers(B, N, H):
    """Canaries behind out / lse / delta / dqkv / probs: every attention kernel (short, key-tile forward, streaming backward,
    tensor-core maps) writes ragged tails (N % 128 != 0) through clipped TMA boxes or guarded stores -- nothing may land past
    the tensors.  (compute-sanitizer is not available on the GPU pool, so the bounds are checked this way.)"""
    pad = 4096

    def guarded(shape, dtype, fill):
        n = 1
        for d in shape:
            n *= d
        flat = torch.full((n + pad,), fill, dtype=dtype, device=DEV)
        return flat, flat[:n].view(*shape)

    qkv = _rand(B, N, 3 * H * 64, dtype=F16, seed=3)
    f_out, out = guarded((B, N, H * 64), F16, 777.0)
    f_lse, lse = guarded((B, H, N), torch.float32, 777.0)
    f_pr, probs = guarded((B, H, N, N), torch.float32, 777.0)
    ops.attention_fwd(qkv, B, N, H, 0.125, out=out, lse=lse, probs=probs)
    dout = _rand(B, N, H * 64, dtype=F16, seed=4)
    f_dq, dqkv = guarded((B, N, 3 * H * 64), F16, 777.0)
    f_de, delta = guarded((B, H, N), torch.float32, 777.0)
    ops.attention_bwd(qkv, out, dout, lse, B, N, H, 0.125, dqkv=dqkv, delta=delta)
    torch.cuda.synchronize()
    for name, flat in (("out", f_out), ("lse", f_lse), ("probs", f_pr), ("dqkv", f_dq), ("delta", f_de)):
        assert (flat[-pad:] == 777.0).all(), name
    for t in (out, lse, probs, dqkv):
        assert torch.isfinite(t.float()).all() and not (t == 777.0).all()
    assert (probs.sum(-1) - 1).abs().max().item() < 1e-5


# ------------------------------------------------------------------ token plumbing
@pytest.mark.parametrize("B,C,S,P", [(3, 3, 224, 16), (2, 1, 256, 16), (2, 3, 64, 8), (1, 3, 64, 32)])
@pytest.mark.parametrize("dt", [F16, BF16])
def test_patchify(B, C, S, P, dt):
    img = _rand(B, C, S, S, seed=1)
    out = ops.patchify(img, P, dtype=dt)
    ref = torch.nn.functional.unfold(img, P, stride=P).transpose(1, 2).reshape(-1, C * P * P)
    torch.cuda.synchronize()
    assert torch.equal(out, ref.to(dt))
    # channel-last patch vectors (PatchEmbed projection_type='linear'): einops 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)'
    out2 = ops.patchify(img, P, dtype=dt, channel_last=True)
    g = S // P
    ref2 = img.reshape(B, C, g, P, g, P).permute(0, 2, 4, 3, 5, 1).reshape(-1, P * P * C)
    torch.cuda.synchronize()
    assert torch.equal(out2, ref2.to(dt))


def test_prefix_and_tokens_bwd():
    B, T, D, npre = 37, 198, 192, 2
    x = torch.zeros(B, T, D, device=DEV)
    cls, dist, pos = _rand(D, seed=1), _rand(D, seed=2), _rand(T, D, seed=3)
    ops.prefix_tokens_fwd(x, cls, dist, pos, npre)
    torch.cuda.synchronize()
    assert torch.allclose(x[:, 0], (cls + pos[0]).expand(B, D)) and torch.allclose(x[:, 1], (dist + pos[1]).expand(B, D))
    assert x[:, 2:].abs().max().item() == 0
    dx = _rand(B, T, D, seed=4)
    dpos = torch.zeros(T, D, device=DEV); dcls = torch.zeros(D, device=DEV); ddist = torch.zeros(D, device=DEV)
    dbias = torch.zeros(D, device=DEV)
    dpatch = torch.empty(B * (T - npre), D, dtype=F16, device=DEV)
    u = torch.tensor([0.5], device=DEV)
    ops.tokens_bwd(dx, dpos, dcls, ddist, dpatch, dbias, npre, unscale=u)
    torch.cuda.synchronize()
    assert rel_l2(dpos, 0.5 * dx.sum(0)) < 1e-5
    assert rel_l2(dcls, 0.5 * dx[:, 0].sum(0)) < 1e-5 and rel_l2(ddist, 0.5 * dx[:, 1].sum(0)) < 1e-5
    assert rel_l2(dbias, 0.5 * dx[:, npre:].sum((0, 1))) < 1e-5
    assert torch.equal(dpatch, dx[:, npre:].reshape(-1, D).to(F16))


# ------------------------------------------------------------------ heads
@pytest.mark.parametrize("n_heads,C", [(2, 2), (1, 2), (1, 10)])
def test_head_fwd_bwd(n_heads, C):
    B, T, D = 19, 198, 192
    x = _rand(B, T, D, seed=1)
    g = _rand(D, seed=2) * 0.1 + 1; b = _rand(D, seed=3) * 0.1
    W0 = _rand(C, D, scale=0.02, seed=4); b0 = _rand(C, seed=5) * 0.1
    W1 = _rand(C, D, scale=0.02, seed=6) if n_heads == 2 else None
    b1 = _rand(C, seed=7) * 0.1 if n_heads == 2 else None
    l0, l1, xhat, rstd = ops.head_fwd(x, g, b, W0, b0, W1, b1, n_heads)
    xr = x.clone().requires_grad_(True)
    params = [t.clone().requires_grad_(True) for t in (g, b, W0, b0)] + ([W1.clone().requires_grad_(True), b1.clone().requires_grad_(True)] if n_heads == 2 else [])
    xn = torch.nn.functional.layer_norm(xr, (D,), params[0], params[1], 1e-5)
    r0 = xn[:, 0] @ params[2].t() + params[3]
    torch.cuda.synchronize()
    assert (l0 - r0).abs().max().item() < 1e-5
    dl0 = _rand(B, C, seed=8); dl1 = _rand(B, C, seed=9) if n_heads == 2 else None
    loss = (r0 * dl0).sum()
    if n_heads == 2:
        r1 = xn[:, 1] @ params[4].t() + params[5]
        assert (l1 - r1).abs().max().item() < 1e-5
        loss = loss + (r1 * dl1).sum()
    loss.backward()
    S = 1024.0
    dx = torch.full((B, T, D), 7.0, device=DEV); dx16 = torch.empty(B, T, D, dtype=F16, device=DEV)
    z = lambda *s: torch.zeros(*s, device=DEV)
    dg, db_, dW0, db0, dcs = z(D), z(D), z(C, D), z(C), z(D)
    dW1, db1 = (z(C, D), z(C)) if n_heads == 2 else (None, None)
    ops.head_bwd(dl0, dl1, xhat, rstd, g, b, W0, W1, dx, dx16, dg, db_, dW0, db0, dW1, db1, dcs, T, n_heads,
                 loss_scale=torch.tensor([S], device=DEV))
    torch.cuda.synchronize()
    assert rel_l2(dx / S, xr.grad) < 1e-5 and dx[:, n_heads:].abs().max().item() == 0
    assert rel_l2(dx16.float() / S, xr.grad) < 1e-3
    assert rel_l2(dg, params[0].grad) < 1e-4 and rel_l2(db_, params[1].grad) < 1e-4
    assert rel_l2(dW0, params[2].grad) < 1e-4 and rel_l2(db0, params[3].grad) < 1e-4
    assert rel_l2(dcs, xr.grad.sum((0, 1))) < 1e-4
    if n_heads == 2:
        assert rel_l2(dW1, params[4].grad) < 1e-4 and rel_l2(db1, params[5].grad) < 1e-4


# ------------------------------------------------------------------ loss
@pytest.mark.parametrize("mode,ls", [(0, 0.0), (0, 0.1), (1, 0.0), (1, 0.1), (2, 0.0)])
def test_loss(mode, ls):
    import torch.nn.functional as F
    B, C, T, alpha = 77, 2, 3.0, 0.7
    cls = _rand(B, C, seed=1); dist = _rand(B, C, seed=2); teacher = _rand(B, C, seed=3) * 2
    y = torch.randint(0, C, (B,), generator=torch.Generator().manual_seed(4)).to(DEV)
    cr = cls.clone().requires_grad_(True); dr = dist.clone().requires_grad_(True)
    if mode == 0:
        w_cls = w_dist = 0.5
        ref = 0.5 * F.cross_entropy(cr, y, label_smoothing=ls) + 0.5 * F.cross_entropy(dr, y, label_smoothing=ls)
        tl = None
    else:
        w_cls, w_dist = 1 - alpha, alpha
        ce = F.cross_entropy(cr, y, label_smoothing=ls)
        if mode == 1:
            dl = F.kl_div(F.log_softmax(dr / T, 1), F.softmax(teacher / T, 1), reduction="batchmean", log_target=False) * T ** 2
        else:
            dl = F.cross_entropy(dr, teacher.argmax(1), label_smoothing=ls)
        ref = (1 - alpha) * ce + alpha * dl
        tl = teacher
    ref.backward()
    out, dcls, ddist = ops.loss_fwd_bwd(cls, dist, tl, y, mode=mode, w_cls=w_cls, w_dist=w_dist, T=T, label_smoothing=ls)
    torch.cuda.synchronize()
    assert abs(out[0].item() - ref.item()) < 1e-5
    assert (dcls - cr.grad).abs().max().item() < 1e-6 and (ddist - dr.grad).abs().max().item() < 1e-6
    assert out[3].item() == (cls.argmax(1) == y).sum().item()
    if tl is not None:
        assert out[4].item() == (cls.argmax(1) == teacher.argmax(1)).sum().item()


# ------------------------------------------------------------------ AdamW + clip + loss-scale bookkeeping
def _flat_setup(sizes):
    offs, total = [], 0
    for s in sizes:
        offs.append(total); total += (math.prod(s) + 127) // 128 * 128
    return offs, total


@pytest.mark.parametrize("max_norm", [0.0, 1.0])
def test_adamw_matches_torch(max_norm):
    sizes = [(192, 768), (192,), (2, 192), (1, 198, 192), (5,)]
    lr, wd_list, scale_list = 1e-3, [0.05, 0.0, 0.05, 0.0, 0.05], [1.0, 0.75, 0.5, 1.0, 0.1]
    offs, total = _flat_setup(sizes)
    flat_p = torch.zeros(total, device=DEV); flat_g = torch.zeros(total, device=DEV)
    m = torch.zeros(total, device=DEV); v = torch.zeros(total, device=DEV)
    p16 = torch.zeros(total, dtype=BF16, device=DEV); ph16 = torch.zeros(total, dtype=F16, device=DEV)
    refs = []
    for i, (s, o) in enumerate(zip(sizes, offs)):
        n = math.prod(s)
        flat_p[o:o + n] = _rand(n, seed=10 + i)
        refs.append(flat_p[o:o + n].clone().view(s).requires_grad_(True))
    opt = torch.optim.AdamW([{"params": [r], "lr": lr * sc, "weight_decay": w} for r, sc, w in zip(refs, scale_list, wd_list)],
                            betas=(0.9, 0.999), eps=1e-8)
    chunk_off, chunk_len, c_scale, c_wd = [], [], [], []
    for s, o, sc, w in zip(sizes, offs, scale_list, wd_list):
        n = (math.prod(s) + 127) // 128 * 128
        for c0 in range(0, n, 4096):
            chunk_off.append(o + c0); chunk_len.append(min(4096, n - c0)); c_scale.append(sc); c_wd.append(w)
    chunk_off = torch.tensor(chunk_off, dtype=torch.int64, device=DEV); chunk_len = torch.tensor(chunk_len, dtype=torch.int32, device=DEV)
    c_scale = torch.tensor(c_scale, device=DEV); c_wd = torch.tensor(c_wd, device=DEV)
    state = torch.tensor([0.0, lr, 0.0, 0.0], device=DEV)
    amp = torch.tensor([1024.0, 1 / 1024.0, 0, 0, 0, 0, 0, 0], device=DEV)
    for step in range(3):
        for i, (s, o) in enumerate(zip(sizes, offs)):
            n = math.prod(s)
            g = _rand(n, seed=100 + 10 * step + i) * 0.3
            flat_g[o:o + n] = g
            refs[i].grad = g.clone().view(s)
        if max_norm > 0:
            torch.nn.utils.clip_grad_norm_(refs, max_norm)
        ops.grad_sqnorm(flat_g, state)
        opt.step()
        ops.adamw_step(flat_p, flat_g, m, v, p16, ph16, chunk_off, chunk_len, c_scale, c_wd, state, amp, 0.9, 0.999, 1e-8,
                       max_norm, growth_interval=2)
    torch.cuda.synchronize()
    for r, s, o in zip(refs, sizes, offs):
        n = math.prod(s)
        assert (flat_p[o:o + n] - r.detach().flatten()).abs().max().item() < 2e-6
        assert torch.equal(p16[o:o + n], flat_p[o:o + n].to(BF16))
        assert torch.equal(ph16[o:o + n], flat_p[o:o + n].to(F16))
    assert state[0].item() == 3.0
    assert amp[0].item() == 2048.0 and amp[2].item() == 1.0      # doubled once after 2 clean steps, 1 clean step since
    # an overflowed step is skipped: parameters untouched, S halved, step counter unchanged
    before = flat_p.clone()
    flat_g[5] = float("inf")
    ops.grad_sqnorm(flat_g, state)
    ops.adamw_step(flat_p, flat_g, m, v, p16, ph16, chunk_off, chunk_len, c_scale, c_wd, state, amp, 0.9, 0.999, 1e-8, max_norm,
                   growth_interval=2)
    torch.cuda.synchronize()
    assert torch.equal(before, flat_p) and state[0].item() == 3.0
    assert amp[0].item() == 1024.0 and amp[3].item() == 1.0 and amp[4].item() == 1.0
    assert abs(amp[1].item() - 1 / 1024.0) < 1e-12


def test_amp_update_external_optimizer_path():
    g = _rand(5000, seed=1)
    amp = torch.tensor([64.0, 1 / 64.0, 0, 0, 0, 0, 0, 0], device=DEV)
    scratch = torch.zeros(4, device=DEV)
    keep = g.clone()
    ops.amp_update(g, amp, scratch, growth_interval=1)
    torch.cuda.synchronize()
    assert torch.equal(g, keep) and amp[0].item() == 128.0
    g[17] = float("nan")
    ops.amp_update(g, amp, scratch, growth_interval=1)
    torch.cuda.synchronize()
    assert g.abs().max().item() == 0 and amp[0].item() == 64.0 and amp[4].item() == 1.0


# ------------------------------------------------------------------ helpers
def test_cast_and_colsum():
    x = _rand(1000, 576, seed=1)
    xb = ops.cast_bf16(x)
    xh = ops.cast_fp16(x)
    assert torch.equal(xb, x.to(BF16)) and torch.equal(xh, x.to(F16))
    for t in (xb, xh):
        out = torch.zeros(576, device=DEV)
        ops.colsum16(t, out, unscale=torch.tensor([0.125], device=DEV))
        torch.cuda.synchronize()
        assert rel_l2(out, 0.125 * t.float().sum(0)) < 1e-5


def test_ensemble_and_rollout():
    F_, B, C = 5, 33, 2
    logits = _rand(F_, B, C, seed=1)
    w = torch.full((F_,), 1.0 / F_, device=DEV)
    probs, pred = ops.ensemble_probs(logits, w)
    ref = (logits.softmax(-1) * w.view(-1, 1, 1)).sum(0)
    torch.cuda.synchronize()
    assert (probs - ref).abs().max().item() < 1e-6
    assert torch.equal(pred, ref.argmax(1))
    L, B, H, N = 3, 2, 3, 50
    p = _rand(L, B, H, N, N, seed=2).softmax(-1).contiguous()
    for fusion in ("mean", "max", "min"):
        r = ops.attention_rollout(p, fusion)
        R = torch.eye(N, device=DEV).expand(B, N, N)
        for l in range(L):
            f = {"mean": p[l].mean(1), "max": p[l].max(1).values, "min": p[l].min(1).values}[fusion]
            a = 0.5 * (f + torch.eye(N, device=DEV))
            a = a / a.sum(-1, keepdim=True)
            R = a @ R
        torch.cuda.synchronize()
        assert (r - R).abs().max().item() < 1e-5
        for row in (0, 7):      # one row by vector-matrix products (the class-token map of config 5): same numbers
            rr = ops.attention_rollout_row(p, row, fusion)
            assert rr.shape == (B, N) and (rr - R[:, row]).abs().max().item() < 1e-5
            rim = ops.attention_rollout_row(p.transpose(0, 1).contiguous(), row, fusion, image_major=True)   # [B,L,H,N,N] layout
            assert torch.equal(rim, rr)


# ------------------------------------------------------------------ nn.Dropout (drop_rate > 0): counter-based masks
def _seed(v: int):
    return torch.tensor([v], dtype=torch.int64, device=DEV)


def test_dropout_mask_is_a_pure_function_of_seed_and_site():
    s1, s2 = _seed(1234567), _seed(1234568)
    m = ops.dropout_mask(s1, 0.1, 3, 5000, 192)
    assert torch.equal(m, ops.dropout_mask(s1, 0.1, 3, 5000, 192))                    # deterministic
    vals = set(m.unique().tolist())
    assert len(vals) == 2 and 0.0 in vals and abs(max(vals) - 1 / 0.9) < 1e-6        # 0 or 1/(1-p)
    assert abs((m == 0).float().mean().item() - 0.1) < 3e-3                          # ~1M draws
    for other in (ops.dropout_mask(s2, 0.1, 3, 5000, 192), ops.dropout_mask(s1, 0.1, 4, 5000, 192)):
        agree = ((other == 0) == (m == 0)).float().mean().item()                     # independent masks agree on 0.82 of elements
        assert abs(agree - 0.82) < 5e-3
    # a row-major [rows, cols] view of the element counter: the same elements for a different column count
    assert torch.equal(m.view(-1), ops.dropout_mask(s1, 0.1, 3, 1250, 768).view(-1))
    # no visible structure along rows or columns
    dropped = (m == 0).float()
    assert (dropped.mean(0) - 0.1).abs().max().item() < 0.03 and (dropped.mean(1) - 0.1).abs().max().item() < 0.12
    assert torch.equal(ops.dropout_mask(s1, 0.0, 3, 64, 64), torch.ones(64, 64, device=DEV))
    for p in (0.05, 0.5, 0.9):
        assert abs((ops.dropout_mask(s1, p, 0, 4096, 256) == 0).float().mean().item() - p) < 3e-3


@pytest.mark.parametrize("M,N,K", [(6336, 192, 192), (1000, 768, 3072)])
def test_gemm_residual_dropout(M, N, K):
    """out = residual + row_scale * mask * (A W^T + bias): proj_drop / Mlp.drop fused into the residual epilogue."""
    A = _rand(M, K, dtype=F16, seed=1)
    W = _rand(N, K, scale=0.05, dtype=F16, seed=2)
    bias, res = _rand(N, seed=3), _rand(M, N, seed=4)
    seed = _seed(99)
    out = torch.empty(M, N, device=DEV)
    ops.gemm(A, W, M, N, K, out=out, bias=bias, residual=res, drop=(seed, 0.1, 7))
    mask = ops.dropout_mask(seed, 0.1, 7, M, N)
    ref = res + mask * (A.float() @ W.float().t() + bias)
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 1e-5
    assert torch.equal(out[mask == 0], res[mask == 0])
    rs = (torch.floor(0.7 + torch.rand(M, device=DEV)) / 0.7)
    ops.gemm(A, W, M, N, K, out=out, bias=bias, residual=res, row_scale=rs, drop=(seed, 0.25, 8))
    ref = res + rs[:, None] * ops.dropout_mask(seed, 0.25, 8, M, N) * (A.float() @ W.float().t() + bias)
    assert rel_l2(out, ref) < 1e-5


@pytest.mark.parametrize("dt", [F16, BF16])
@pytest.mark.parametrize("M,N,K", [(6336, 768, 192), (777, 3072, 768)])
def test_gemm_gelu_dropout(M, N, K, dt):
    """Mlp.drop after the activation: out2 = mask * gelu(pre) and the saved derivative out = mask * gelu'(pre)."""
    A = _rand(M, K, dtype=dt, seed=1)
    W = _rand(N, K, scale=0.1, dtype=dt, seed=2)
    bias = _rand(N, seed=3)
    seed = _seed(5)
    dact = torch.empty(M, N, dtype=dt, device=DEV)
    act = torch.empty(M, N, dtype=dt, device=DEV)
    ops.gemm(A, W, M, N, K, out=dact, out2=act, bias=bias, epilogue=_lib.EPI_GELU, drop=(seed, 0.1, 2))
    mask = ops.dropout_mask(seed, 0.1, 2, M, N)
    pre = (A.float() @ W.float().t() + bias).requires_grad_(True)
    ref_act = torch.nn.functional.gelu(pre) * mask
    ref_act.sum().backward()
    torch.cuda.synchronize()
    assert rel_l2(act, ref_act) < OUT_TOL[dt] and rel_l2(dact, pre.grad) < OUT_TOL[dt]
    assert act[mask == 0].abs().max().item() == 0 and dact[mask == 0].abs().max().item() == 0


def test_tokens_dropout_fwd_bwd():
    """pos_drop over the assembled token matrix: patch rows in the GEMM epilogue, cls/dist rows in the prefix kernel,
    and the same mask on the gradient in the token-assembly backward."""
    B, P, T, prefix, D, K = 5, 196, 198, 2, 192, 768
    A = _rand(B * P, K, dtype=F16, seed=1)
    W = _rand(D, K, scale=0.05, dtype=F16, seed=2)
    bias, pos = _rand(D, seed=3), _rand(T, D, seed=4)
    cls, dist = _rand(D, seed=5), _rand(D, seed=6)
    seed = _seed(31)
    x = torch.zeros(B, T, D, dtype=torch.float32, device=DEV)
    ops.gemm(A, W, B * P, D, K, out=x, bias=bias, epilogue=_lib.EPI_TOKENS, tokens=(P, T, prefix), pos=pos, drop=(seed, 0.2, 0))
    ops.prefix_tokens_fwd(x, cls, dist, pos, prefix, drop=(seed, 0.2, 0))
    mask = ops.dropout_mask(seed, 0.2, 0, B * T, D).view(B, T, D)
    ref = torch.empty_like(x)
    ref[:, prefix:] = (A.float() @ W.float().t() + bias).view(B, P, D) + pos[prefix:]
    ref[:, 0] = cls + pos[0]
    ref[:, 1] = dist + pos[1]
    ref = ref * mask
    torch.cuda.synchronize()
    assert (x - ref).abs().max().item() < 1e-3 and x[mask == 0].abs().max().item() == 0
    dx = _rand(B, T, D, seed=7)
    z = lambda *s: torch.zeros(*s, device=DEV)
    dpos, dcls, ddist, dbias = z(T, D), z(D), z(D), z(D)
    dpatch = torch.empty(B * P, D, dtype=F16, device=DEV)
    ops.tokens_bwd(dx, dpos, dcls, ddist, dpatch, dbias, prefix, drop=(seed, 0.2, 0))
    g = dx * mask
    torch.cuda.synchronize()
    assert rel_l2(dpos, g.sum(0)) < 1e-5 and rel_l2(dcls, g[:, 0].sum(0)) < 1e-5 and rel_l2(ddist, g[:, 1].sum(0)) < 1e-5
    assert rel_l2(dbias, g[:, prefix:].sum((0, 1))) < 1e-5
    assert rel_l2(dpatch.float(), g[:, prefix:].reshape(B * P, D)) < 1e-3


@pytest.mark.parametrize("rows,dim", [(6336, 192), (197 * 3, 768), (333, 256)])
def test_layernorm_bwd_branch_dropout(rows, dim):
    """dx16 / dcolsum = dx * branch_scale * dropout mask of the branch below; the fp32 residual-stream dx is untouched."""
    x = _rand(rows, dim, seed=1)
    gamma = _rand(dim, seed=2) * 0.1 + 1
    beta = _rand(dim, seed=3) * 0.1
    y, mean, rstd = ops.layernorm_fwd(x, gamma, beta)
    dy = _rand(rows, dim, dtype=F16, seed=4)
    dres = _rand(rows, dim, seed=5)
    z = lambda *s: torch.zeros(*s, device=DEV)
    seed = _seed(77)
    outs = []
    for drop in (None, (seed, 0.1, 11)):
        dg, db, dcs = z(dim), z(dim), z(dim)
        dx16 = torch.empty(rows, dim, dtype=F16, device=DEV)
        bs = torch.floor(0.8 + torch.rand(rows, generator=torch.Generator().manual_seed(1))).to(DEV) / 0.8
        dx = ops.layernorm_bwd(dy, x, mean, rstd, gamma, dg, db, dres=dres, dx16=dx16, dcolsum=dcs, branch_scale=bs,
                               branch_drop=drop)
        outs.append((dx.clone(), dx16.float(), dcs, dg, db, bs))
    mask = ops.dropout_mask(seed, 0.1, 11, rows, dim)
    (dx0, h0, c0, g0, b0, bs), (dx1, h1, c1, g1, b1, _) = outs
    torch.cuda.synchronize()
    assert rel_l2(g0, g1) < 1e-5 and rel_l2(b0, b1) < 1e-5       # dgamma / dbeta do not see the branch mask
    assert torch.equal(dx0, dx1)
    ref16 = dx0 * bs[:, None] * mask
    assert rel_l2(h1, ref16) < 1e-3 and h1[mask == 0].abs().max().item() == 0
    assert rel_l2(c1, ref16.sum(0)) < 1e-4


def test_head_bwd_branch_dropout():
    B, T, D, C = 19, 198, 192, 2
    x = _rand(B, T, D, seed=1)
    g = _rand(D, seed=2) * 0.1 + 1; b = _rand(D, seed=3) * 0.1
    W0 = _rand(C, D, scale=0.02, seed=4); b0 = _rand(C, seed=5) * 0.1
    l0, _, xhat, rstd = ops.head_fwd(x, g, b, W0, b0, None, None, 1)
    dl0 = _rand(B, C, seed=8)
    z = lambda *s: torch.zeros(*s, device=DEV)
    seed = _seed(3)
    res = []
    for drop in (None, (seed, 0.3, 6)):
        dx = torch.empty(B, T, D, device=DEV); dx16 = torch.empty(B, T, D, dtype=F16, device=DEV)
        dcs = z(D)
        ops.head_bwd(dl0, None, xhat, rstd, g, b, W0, None, dx, dx16, z(D), z(D), z(C, D), z(C), None, None, dcs, T, 1,
                     branch_drop=drop)
        res.append((dx.clone(), dx16.float(), dcs))
    mask = ops.dropout_mask(seed, 0.3, 6, B * T, D).view(B, T, D)
    torch.cuda.synchronize()
    assert torch.equal(res[0][0], res[1][0])
    assert rel_l2(res[1][1], res[0][0] * mask) < 1e-3
    assert rel_l2(res[1][2], (res[0][0] * mask).sum((0, 1))) < 1e-4


# ------------------------------------------------------------------ general classification tail (gap pooling, pre_logits)
@pytest.mark.parametrize("B,T,D,t0,t1", [(19, 197, 192, 1, 197), (5, 196, 768, 0, 196), (3, 17, 64, 0, 1), (2, 50, 384, 7, 31)])
@pytest.mark.parametrize("dt", [F16, BF16])
def test_pool_norm_fwd_bwd(B, T, D, t0, t1, dt):
    """vitk_pool_norm_{fwd,bwd} against torch fp32: mean over tokens [t0,t1) of LayerNorm(x) (vision_transformer_base.py:469-474).
    fp32 kernels: 1e-5; the 16-bit gradient copy: rounding of the format."""
    x = _rand(B, T, D, seed=1)
    g = _rand(D, seed=2) * 0.1 + 1; b = _rand(D, seed=3) * 0.1
    pooled, mean, rstd = ops.pool_norm_fwd(x, g, b, t0, t1)
    xr = x.clone().requires_grad_(True); gr = g.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-5)[:, t0:t1].mean(1)
    torch.cuda.synchronize()
    assert (pooled - ref).abs().max().item() < 1e-5
    dp = _rand(B, D, seed=4)
    (ref * dp).sum().backward()
    S = 512.0
    bs = (torch.rand(B * T, device=DEV) > 0.3).float() / 0.7            # stochastic-depth factor of the last MLP branch
    dx = torch.full((B, T, D), 7.0, device=DEV); dx16 = torch.full((B, T, D), 3.0, dtype=dt, device=DEV)
    dg, db_, dcs = (torch.zeros(D, device=DEV) for _ in range(3))
    ops.pool_norm_bwd(dp, x, mean, rstd, g, dx, dx16, dg, db_, dcs, t0, t1, loss_scale=torch.tensor([S], device=DEV),
                      branch_scale=bs)
    torch.cuda.synchronize()
    assert rel_l2(dx / S, xr.grad) < 1e-5
    outside = torch.ones(T, dtype=torch.bool); outside[t0:t1] = False
    assert dx[:, outside].abs().sum().item() == 0 and dx16[:, outside].float().abs().sum().item() == 0
    branch = xr.grad * bs.view(B, T, 1)
    assert rel_l2(dx16.float() / S, branch) < OUT_TOL[dt] * 2
    assert rel_l2(dg, gr.grad) < 1e-4 and rel_l2(db_, br.grad) < 1e-4
    assert rel_l2(dcs, branch.sum((0, 1))) < 1e-4


def test_pool_norm_bwd_branch_dropout():
    B, T, D = 7, 65, 128
    x = _rand(B, T, D, seed=1)
    g = _rand(D, seed=2) * 0.1 + 1; b = _rand(D, seed=3) * 0.1
    _, mean, rstd = ops.pool_norm_fwd(x, g, b, 1, T)
    dp = _rand(B, D, seed=4)
    seed = _seed(5)
    res = []
    for drop in (None, (seed, 0.25, 4)):
        dx = torch.empty(B, T, D, device=DEV); dx16 = torch.empty(B, T, D, dtype=F16, device=DEV)
        dcs = torch.zeros(D, device=DEV)
        ops.pool_norm_bwd(dp, x, mean, rstd, g, dx, dx16, torch.zeros(D, device=DEV), torch.zeros(D, device=DEV), dcs, 1, T,
                          branch_drop=drop)
        res.append((dx.clone(), dx16.float(), dcs))
    mask = ops.dropout_mask(seed, 0.25, 4, B * T, D).view(B, T, D)
    torch.cuda.synchronize()
    assert torch.equal(res[0][0], res[1][0])                 # the residual path is not masked
    assert rel_l2(res[1][1], res[0][0] * mask) < 1e-3
    assert rel_l2(res[1][2], (res[0][0] * mask).sum((0, 1))) < 1e-4


@pytest.mark.parametrize("B,I,O_,act", [(19, 192, 192, 1), (256, 768, 768, 1), (33, 128, 2, 0), (5, 64, 10, 0)])
def test_dense_fwd_bwd(B, I, O_, act):
    """vitk_dense_{fwd,bwd} (pre_logits Linear+Tanh :380-386 and the head Linear) against torch fp32, gradients ACCUMULATE."""
    x = _rand(B, I, seed=1); W = _rand(O_, I, scale=0.05, seed=2); bias = _rand(O_, seed=3) * 0.1
    y = ops.dense_fwd(x, W, bias, act=act)
    xr, Wr, br = (t.clone().requires_grad_(True) for t in (x, W, bias))
    ref = xr @ Wr.t() + br
    ref = torch.tanh(ref) if act == 1 else ref
    torch.cuda.synchronize()
    assert (y - ref).abs().max().item() < 1e-5           # fp32 dot products of up to 768 terms, different summation order
    dy = _rand(B, O_, seed=4)
    (ref * dy).sum().backward()
    dW = torch.ones(O_, I, device=DEV); db = torch.ones(O_, device=DEV)
    dx = ops.dense_bwd(dy, y if act == 1 else None, x, W, dW, db, act=act)
    torch.cuda.synchronize()
    assert rel_l2(dx, xr.grad) < 1e-5
    assert rel_l2(dW - 1, Wr.grad) < 1e-5 and rel_l2(db - 1, br.grad) < 1e-5
    assert ops.dense_bwd(dy, y if act == 1 else None, x, W, dW, None, act=act, need_dx=False) is None


def test_general_tail_rejects_bad_arguments():
    x = _rand(2, 9, 64, seed=1); g = torch.ones(64, device=DEV); b = torch.zeros(64, device=DEV)
    with pytest.raises(RuntimeError):
        ops.pool_norm_fwd(x, g, b, 3, 3)             # empty token range
    with pytest.raises(RuntimeError):
        ops.pool_norm_fwd(x, g, b, 0, 10)            # beyond the sequence
    with pytest.raises(RuntimeError):
        ops.dense_fwd(x[:, 0], torch.ones(4, 64, device=DEV), None, act=2)


# ------------------------------------------------------------------ on-device validation / test metrics
@pytest.mark.parametrize("quant", [None, 4])
def test_metrics_counters_and_auroc_match_oracle(quant):
    """Confusion counters and the pairwise AUROC against the oracle's restatement of torchmetrics (pinned to scikit-learn
    in tests/test_oracle.py); batches of ragged sizes are appended across several update launches."""
    from oracle import vit_oracle as O
    from thyroid_vit_cnn_comparison_b200.metrics import ClassificationMetrics
    g = torch.Generator().manual_seed(11)
    met = ClassificationMetrics(2, capacity=8192)
    all_logits, all_y = [], []
    for B in (256, 37, 1, 450, 1000, 3):
        logits = torch.randn(B, 2, generator=g) * 2
        if quant is not None:
            logits = torch.round(logits * quant) / quant             # tied scores: the ties/2 term matters
        y = torch.randint(0, 2, (B,), generator=g)
        met.update(logits.to(DEV), y.to(DEV))
        all_logits.append(logits); all_y.append(y)
    logits, y = torch.cat(all_logits), torch.cat(all_y)
    got = met.compute()
    tp, fp, tn, fn = O.binary_stat_scores(logits.argmax(1), y)
    ref = O.binary_metrics(tp, fp, tn, fn)
    assert got["stat_scores"] == ref["stat_scores"]                  # integer counters: exact
    for k in ("acc", "f1", "specificity", "sensitivity", "ppv", "npv"):
        assert got[k] == ref[k], k
    ref_auc = O.binary_auroc(torch.softmax(logits, 1)[:, 1], y)
    assert abs(got["auc"] - ref_auc) < (1e-12 if quant is not None else 2e-6), (got["auc"], ref_auc)
    assert int(met.count.item()) == logits.shape[0]
    met.reset()
    met.update(logits[:5].to(DEV), torch.ones(5, dtype=torch.long, device=DEV))
    assert met.compute()["auc"] == 0.0                               # a class is absent
    met.reset()
    met.update(logits[:4].to(DEV), torch.tensor([0, 1, 2, 1], device=DEV))
    with pytest.raises(ValueError):
        met.compute()                                                # label outside [0, C)
    small = ClassificationMetrics(2, capacity=16)
    small.update(logits[:32].to(DEV), y[:32].to(DEV))
    with pytest.raises(RuntimeError):
        small.compute()                                              # AUROC buffer overflow is loud


def test_metrics_large_split_auroc_and_multiclass_confusion():
    from oracle import vit_oracle as O
    from thyroid_vit_cnn_comparison_b200.metrics import ClassificationMetrics
    g = torch.Generator().manual_seed(5)
    n = 50000
    y = torch.randint(0, 2, (n,), generator=g)
    logits = torch.randn(n, 2, generator=g) + torch.stack([-0.5 * y.float(), 0.5 * y.float()], 1)   # informative scores
    met = ClassificationMetrics(2, capacity=1 << 16)
    for i in range(0, n, 4096):
        met.update(logits[i:i + 4096].to(DEV), y[i:i + 4096].to(DEV))
    got = met.compute()
    ref_auc = O.binary_auroc(torch.softmax(logits, 1)[:, 1], y)
    assert 0.6 < ref_auc < 0.9 and abs(got["auc"] - ref_auc) < 2e-6
    mc = ClassificationMetrics(5)
    lg = torch.randn(999, 5, generator=g)
    yy = torch.randint(0, 5, (999,), generator=g)
    mc.update(lg.to(DEV), yy.to(DEV))
    cm = mc.compute()["confusion"]
    ref = torch.zeros(5, 5, dtype=torch.int64)
    for t, p in zip(yy.tolist(), lg.argmax(1).tolist()):
        ref[t, p] += 1
    assert torch.equal(cm, ref)


# ------------------------------------------------------------------ GPU-side input pipeline
def _u16_cuda(a):
    import numpy as np
    return torch.from_numpy(a.view(np.int16).copy()).view(torch.uint16).to(DEV)


def test_ingest_resize_matches_oracle_and_reference_fixture():
    """Bit-exact against the oracle's restatement of cv2's INTER_LINEAR loop (same fp32 operations, no FMA contraction);
    within 2 units of the uint16 grid of the reference's own cv2 output stored in tests/golden/ingest.pt."""
    import numpy as np
    from pathlib import Path
    from oracle import ingest_oracle as IO
    rec = torch.load(Path(__file__).parent / "golden" / "ingest.pt", weights_only=False)
    pre = rec["preprocessed_u16"].numpy().view(np.uint16).astype(np.int64)
    for i, raw in enumerate(rec["raw"]):
        img = raw.numpy().view(np.uint16)
        got = ops.resize_u16(_u16_cuda(img)[None], 224, 224).cpu()[0]
        assert torch.equal(got, IO.preprocess_image(img, 224)[0])
        k = torch.round(got * 65535.0).numpy().astype(np.int64)
        assert np.abs(k - pre[i, 0]).max() <= 2
    # a ragged batch of identical-size tiles, up- and down-scaling, odd sizes
    rng = np.random.default_rng(3)
    for (hs, ws, h) in [(61, 450, 224), (1000, 333, 256), (224, 224, 224), (7, 9, 64)]:
        raw = rng.integers(0, 65536, (3, hs, ws)).astype(np.uint16)
        got = ops.resize_u16(_u16_cuda(raw), h, h).cpu()
        ref = torch.stack([IO.preprocess_image(raw[b], h)[0] for b in range(3)])
        assert torch.equal(got, ref), (hs, ws, h)


@pytest.mark.parametrize("n,B", [(224 * 224, 5), (1000, 3), (7, 2), (65536, 2)])
def test_percentile_bounds_match_torch_quantile(n, B):
    from oracle import ingest_oracle as IO
    g = torch.Generator().manual_seed(n)
    x = torch.rand(B, n, generator=g) ** 2
    x[0, : n // 3] = 0.25                                  # heavy ties around a quantile
    if B > 1:
        x[1] = torch.round(x[1] * 50) / 50                 # few distinct values
    for q in [(1, 99), (0, 100), (50, 50), (12.5, 87.5)]:
        got = ops.percentile_bounds(x.to(DEV), q[0] / 100, q[1] / 100).cpu()
        ref = IO.percentile_bounds(x, q)
        assert (got - ref).abs().max().item() <= 1e-6, (q, got, ref)
    neg = torch.randn(2, 999, generator=g)                 # negative values: the key transform must keep the order
    assert (ops.percentile_bounds(neg.to(DEV), 0.01, 0.99).cpu() - IO.percentile_bounds(neg, (1, 99))).abs().max().item() <= 1e-6


def test_finish_tiles_pipeline_and_mixing_match_reference_fixture():
    import numpy as np
    from pathlib import Path
    from oracle import ingest_oracle as IO
    from thyroid_vit_cnn_comparison_b200 import ingest as ING
    rec = torch.load(Path(__file__).parent / "golden" / "ingest.pt", weights_only=False)
    pre = rec["preprocessed_u16"].numpy().view(np.uint16)
    batch = torch.from_numpy(pre.astype(np.float32) / np.float32(65535.0))                # [4,1,224,224]: the reference's own output
    # AdaptiveNormalization -> 3 channels -> Normalize, against the reference's recorded `adaptive`
    bounds = ops.percentile_bounds(batch.to(DEV), 0.01, 0.99)
    got = ops.finish_tiles(batch[:, 0].contiguous().to(DEV), 3, bounds=bounds, mean=ING.IMAGENET_MEAN, std=ING.IMAGENET_STD).cpu()
    ref = IO.to_channels_and_normalize(rec["adaptive"], 3, ING.IMAGENET_MEAN, ING.IMAGENET_STD)
    assert (got - ref).abs().max().item() < 2e-6
    plain = ops.finish_tiles(batch[:, 0].contiguous().to(DEV), 1).cpu()
    assert torch.equal(plain, batch)
    # MixUp / CutMix with the reference's recorded host draws
    images = torch.randn(4, 3, 64, 64, generator=torch.Generator().manual_seed(rec["mix_seed"]))
    m, c = rec["mixup"], rec["cutmix"]
    assert torch.equal(ING._mix_planes(images.to(DEV), m["index"], lam=m["lam"]).cpu(), m["out"])
    box = IO.rand_bbox(images.shape, c["lam_drawn"], c["cx"], c["cy"])
    assert torch.equal(ING._mix_planes(images.to(DEV), c["index"], cutmix=True, box=box).cpu(), c["out"])
    # the drop-in classes replay the same host RNG stream as the reference's classes
    np.random.seed(123); torch.manual_seed(123)
    mixed, la, lb, lam = ING.MixUp(alpha=0.8)(images.to(DEV), rec["labels"].to(DEV))
    assert torch.equal(mixed.cpu(), m["out"]) and lam == m["lam"] and torch.equal(lb.cpu(), rec["labels"][m["index"]])
    np.random.seed(321); torch.manual_seed(321)
    cut, la, lb, lam = ING.CutMix(alpha=1.0)(images.to(DEV), rec["labels"].to(DEV))
    assert torch.equal(cut.cpu(), c["out"]) and abs(lam - c["lam"]) < 1e-12
    # whole pipeline object on raw tiles, MixUp folded into the last pass
    raw = rec["raw"][1].numpy().view(np.uint16)                                          # 224 x 224: no resize ambiguity
    tiles = np.stack([raw, raw[::-1].copy(), raw[:, ::-1].copy()])
    ing = ING.TileIngest(224, 3, percentiles=(1, 99))
    perm = torch.tensor([2, 0, 1])
    out = ing(_u16_cuda(tiles), mix={"perm": perm, "lam": 0.3}).cpu()
    x = torch.stack([IO.preprocess_image(t, 224) for t in tiles])
    x = IO.to_channels_and_normalize(IO.adaptive_normalization(x), 3, ING.IMAGENET_MEAN, ING.IMAGENET_STD)
    assert (out - IO.mixup(x, perm, 0.3)).abs().max().item() < 5e-6


# ------------------------------------------------------------------ frozen-teacher fast path
@pytest.mark.parametrize("pixels,Ct,C,Co", [(2 * 56 * 56, 256, 64, 64), (3 * 14 * 14, 1280, 1248, 1248), (5 * 7 * 7, 1664, 1664, 1664),
                                            (100, 96, 32, 64)])
@pytest.mark.parametrize("dt", [F16, BF16])
@pytest.mark.parametrize("relu", [True, False])
def test_affine_relu_nhwc(pixels, Ct, C, Co, dt, relu):
    """vitk_affine_relu_nhwc: eval BatchNorm (+ ReLU) over the first C channels of an NHWC buffer with pixel pitch Ct, written with
    pixel pitch Co -- against torch fp32 rounded once to the 16-bit format (bit-exact: one fma + one rounding per element)."""
    x = _rand(pixels, Ct, dtype=dt, seed=1)
    scale = _rand(C, seed=2) * 0.3 + 1.0
    shift = _rand(C, seed=3) * 0.5
    out = torch.full((pixels, Co), 7.0, dtype=dt, device=DEV)
    ops.affine_relu_nhwc(x, C, scale, shift, out=out, relu=relu)
    ref = torch.addcmul(shift.double(), x[:, :C].double(), scale.double())
    ref = (ref.clamp_min(0) if relu else ref)
    torch.cuda.synchronize()
    got = out[:, :C].double()
    ulp = 2.0 ** -10 if dt == F16 else 2.0 ** -7
    assert ((got - ref).abs() <= ulp * ref.abs().clamp_min(1e-3)).all()
    if Co > C:
        assert (out[:, C:] == 7.0).all()                      # channels beyond C are not touched
    y = ops.affine_relu_nhwc(x.view(1, pixels, 1, Ct), C, scale, shift, relu=relu)
    assert y.shape == (1, pixels, 1, C) and torch.equal(y.view(pixels, C), out[:, :C])
    with pytest.raises(RuntimeError):
        ops.affine_relu_nhwc(x, C - 4, scale, shift)          # C must be a multiple of 8


@pytest.mark.parametrize("B,H,W,C,Ct,k,s,p,is_max", [(2, 112, 112, 64, 256, 3, 2, 1, True), (2, 56, 56, 128, 512, 2, 2, 0, False),
                                                     (3, 14, 14, 640, 1664, 2, 2, 0, False), (1, 9, 7, 8, 8, 3, 2, 1, True)])
@pytest.mark.parametrize("dt", [F16, BF16])
def test_pool_nhwc(B, H, W, C, Ct, k, s, p, is_max, dt):
    """vitk_pool_nhwc against torch's pooling of the same 16-bit tensor (max: bit-exact; average: one rounding of the exact
    mean), stored with the block buffer's pitch."""
    import torch.nn.functional as F
    x = _rand(B, H, W, C, dtype=dt, seed=1)
    OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    out = torch.full((B, OH, OW, Ct), 5.0, dtype=dt, device=DEV)
    ops.pool_nhwc(x, out, k, s, p, is_max)
    xn = x.permute(0, 3, 1, 2)
    torch.cuda.synchronize()
    if is_max:
        assert torch.equal(out[..., :C], F.max_pool2d(xn, k, s, p).permute(0, 2, 3, 1))
    else:                                                     # fp32 sum, one rounding to the 16-bit format
        ref = F.avg_pool2d(xn.double(), k, s, p).permute(0, 2, 3, 1)
        ulp = 2.0 ** -10 if dt == F16 else 2.0 ** -7
        assert ((out[..., :C].double() - ref).abs() <= ulp * ref.abs().clamp_min(1e-3)).all()
    if Ct > C:
        assert (out[..., C:] == 5.0).all()
    with pytest.raises(RuntimeError):
        ops.pool_nhwc(x, out[:, :-1].contiguous(), k, s, p, is_max)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("P,Ct,C", [(1000, 256, 64), (128 * 3 + 5, 512, 96), (4096, 1664, 1632), (70, 64, 64), (20000, 224, 224)])
def test_dense_bottleneck_matches_affine_relu_conv1x1(P, Ct, C, dt):
    """vitk_dense_bottleneck = relu(relu(x[:, :C] * s + h) @ W^T + b) with the affine map in packed 16-bit fma.rn.relu (scale and
    shift rounded to the operand type, the product-sum rounded once): against that arithmetic restated in torch followed by an
    fp32 matmul, and against the fp32-parameter form (vitk_affine_relu_nhwc) within the rounding of s and h; ragged pixel
    counts, C not a multiple of 64, pitch > C."""
    g = torch.Generator().manual_seed(P + C)
    x = (torch.randn(P, Ct, generator=g)).to(DEV).to(dt)
    s = (torch.rand(C, generator=g) + 0.5).to(DEV)
    h = (torch.randn(C, generator=g) * 0.3).to(DEV)
    w = (torch.randn(128, C, generator=g) / C ** 0.5).to(DEV).to(dt)
    b = (torch.randn(128, generator=g) * 0.2).to(DEV)
    out = ops.dense_bottleneck(x, C, s, h, w, b)
    a = torch.relu((x[:, :C].float() * s.to(dt).float() + h.to(dt).float()).to(dt))    # products of two 16-bit values are exact in fp32
    ref = torch.relu(a.float() @ w.float().t() + b)
    a32 = ops.affine_relu_nhwc(x.view(1, 1, P, Ct), C, s, h).view(P, C)              # fp32 scale / shift, one rounding
    ref32 = torch.relu(a32.float() @ w.float().t() + b)
    torch.cuda.synchronize()
    assert out.shape == (P, 128) and out.dtype == dt
    err = (out.float() - ref).abs().max().item()
    ulp = 2.0 ** -8 if dt == torch.bfloat16 else 2.0 ** -11
    tol = ulp * max(1.0, ref.abs().max().item()) * 1.5
    assert err <= tol, (err, tol)
    assert (out.float() - ref32).abs().max().item() <= 4 * tol                        # rounding s and h costs a few output ulps at most
    assert torch.equal(ops.dense_bottleneck(x, C, s, h, w, b), out)          # deterministic
    with pytest.raises(TypeError):
        ops.dense_bottleneck(x, C, s, h, w[:, :-8].contiguous(), b)


def test_im2col_rows_plus_gemm_is_the_stem_convolution():
    """vitk_im2col_rows (7x7 / stride 2 / pad 3 over a 3-channel NHWC tensor, element order ky, kx, c) + vitk_gemm against
    the filters reshaped the same way == F.conv2d."""
    g = torch.Generator().manual_seed(4)
    B, H, W, C, Co = 3, 38, 42, 3, 64
    x = torch.randn(B, C, H, W, generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn(Co, C, 7, 7, generator=g) * 0.1).to(DEV).to(torch.bfloat16)
    b = torch.randn(Co, generator=g).to(DEV)
    xn = x.permute(0, 2, 3, 1).contiguous()
    patches = ops.im2col_rows(xn, 7, 2, 3)
    OH, OW = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    assert patches.shape == (B * OH * OW, 152)
    ref_p = torch.nn.functional.unfold(x.float(), 7, padding=3, stride=2)                    # [B, C*49, L], order (c, ky, kx)
    ref_p = ref_p.view(B, C, 49, OH * OW).permute(0, 3, 2, 1).reshape(B * OH * OW, 147)      # -> (ky*7+kx, c)
    torch.cuda.synchronize()
    assert torch.equal(patches[:, :147].float(), ref_p) and patches[:, 147:].abs().max().item() == 0
    wm = torch.nn.functional.pad(w.permute(0, 2, 3, 1).reshape(Co, 147), (0, 5)).contiguous()
    out = torch.empty(B * OH * OW, Co, dtype=torch.bfloat16, device=DEV)
    ops.gemm(patches, wm, patches.shape[0], Co, 152, out=out, bias=b)
    ref = torch.nn.functional.conv2d(x.float(), w.float(), b, stride=2, padding=3).permute(0, 2, 3, 1).reshape(-1, Co)
    torch.cuda.synchronize()
    assert (out.float() - ref).abs().max().item() < 2.0 ** -7 * max(1.0, ref.abs().max().item())


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,dt,relu", [(3, 224, 224, torch.float16, True), (2, 64, 64, torch.bfloat16, True),
                                          (1, 40, 72, torch.float16, False), (9, 224, 224, torch.bfloat16, True)])
def test_stem_conv7_matches_conv2d(B, H, W, dt, relu):
    """vitk_stem_conv7 (implicit GEMM, zero padding by TMA fill, ragged tiles clipped) vs F.conv2d on the same 16-bit operands
    (torchvision densenet.py features.conv0 + folded norm0 + relu0)."""
    g = torch.Generator().manual_seed(B * H + W)
    x = torch.randn(B, 3, H, W, generator=g).cuda().to(dt)
    w = (torch.randn(64, 3, 7, 7, generator=g) * 0.1).cuda().to(dt)
    bias = torch.randn(64, generator=g).cuda()
    ref = torch.nn.functional.conv2d(x.float(), w.float(), bias, stride=2, padding=3)
    if relu:
        ref = ref.relu()
    xn = x.permute(0, 2, 3, 1).contiguous()
    out = ops.stem_conv7(xn, ops.stem_conv7_weights(w, dt), bias, relu=relu)
    torch.cuda.synchronize()
    assert out.shape == (B, H // 2, W // 2, 64)
    got = out.permute(0, 3, 1, 2).float()
    tol = 2e-2 if dt == torch.bfloat16 else 3e-3       # one rounding of the 16-bit output (|values| up to ~4)
    assert (got - ref).abs().max().item() < tol * max(1.0, ref.abs().max().item()), (got - ref).abs().max().item()
    assert rel_l2(got, ref) < (4e-3 if dt == torch.bfloat16 else 5e-4)
