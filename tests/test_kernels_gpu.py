"""Kernel-level parity: every C-ABI entry point of libvitk.so against a plain PyTorch fp32
reference of the same op, on the GPU.  Tolerances are stated per test (bf16 operands, fp32
accumulation)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from thyroid_vit_cnn_comparison_b200 import _lib, ops  # noqa: E402

DEV = "cuda"


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
def test_gemm_tn_bias(M, N, K):
    A = _rand(M, K, dtype=torch.bfloat16, seed=1)
    W = _rand(N, K, scale=0.05, dtype=torch.bfloat16, seed=2)
    bias = _rand(N, seed=3)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(A, W, M, N, K, out=out, bias=bias)
    ref = A.float() @ W.float().t() + bias
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 6e-3  # bf16 output rounding


@pytest.mark.parametrize("M,N,K", [(6336, 192, 192), (1000, 192, 768), (130, 136, 72)])
def test_gemm_residual_fp32(M, N, K):
    A = _rand(M, K, dtype=torch.bfloat16, seed=1)
    W = _rand(N, K, scale=0.05, dtype=torch.bfloat16, seed=2)
    bias = _rand(N, seed=3)
    res = _rand(M, N, seed=4)
    out = torch.empty(M, N, dtype=torch.float32, device=DEV)
    ops.gemm(A, W, M, N, K, out=out, bias=bias, residual=res)
    ref = A.float() @ W.float().t() + bias + res
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 1e-5
    assert (out - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("M,N,K", [(6336, 768, 192), (300, 256, 128)])
def test_gemm_gelu(M, N, K):
    A = _rand(M, K, dtype=torch.bfloat16, seed=1)
    W = _rand(N, K, scale=0.1, dtype=torch.bfloat16, seed=2)
    bias = _rand(N, seed=3)
    pre = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    act = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(A, W, M, N, K, out=pre, out2=act, bias=bias, epilogue=_lib.EPI_GELU)
    ref = A.float() @ W.float().t() + bias
    torch.cuda.synchronize()
    assert rel_l2(pre, ref) < 6e-3
    assert rel_l2(act, torch.nn.functional.gelu(ref)) < 6e-3
    # fp16 operands / outputs with the bf16 twin of the activation (the training forward's configuration)
    Ah, Wh = A.to(torch.float16), W.to(torch.float16)
    preh = torch.empty(M, N, dtype=torch.float16, device=DEV); acth = torch.empty_like(preh); twin = torch.empty_like(act)
    ops.gemm(Ah, Wh, M, N, K, out=preh, out2=acth, out3=twin, bias=bias, epilogue=_lib.EPI_GELU)
    refh = Ah.float() @ Wh.float().t() + bias
    torch.cuda.synchronize()
    assert rel_l2(preh, refh) < 6e-4 and rel_l2(acth, torch.nn.functional.gelu(refh)) < 6e-4
    assert rel_l2(twin, torch.nn.functional.gelu(refh)) < 6e-3


@pytest.mark.parametrize("M,N,K", [(6336, 768, 192), (6336, 192, 576), (333, 3072, 768), (130, 72, 136)])
def test_gemm_dgrad_b_mn_major(M, N, K):
    """dX[M,N] = dY[M,K] @ W[K,N]  -- W read MN-major exactly as stored by nn.Linear ([out=K, in=N])."""
    dY = _rand(M, K, dtype=torch.bfloat16, seed=1)
    W = _rand(K, N, scale=0.05, dtype=torch.bfloat16, seed=2)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(dY, W, M, N, K, b_mn=True, out=out)
    ref = dY.float() @ W.float()
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 6e-3


def test_gemm_dgelu():
    M, N, K = 1000, 768, 192
    dY = _rand(M, K, dtype=torch.bfloat16, seed=1)
    W = _rand(K, N, scale=0.05, dtype=torch.bfloat16, seed=2)
    pre = _rand(M, N, dtype=torch.bfloat16, seed=5)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(dY, W, M, N, K, b_mn=True, out=out, aux=pre, epilogue=_lib.EPI_DGELU)
    x = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(x).backward(dY.float() @ W.float())
    torch.cuda.synchronize()
    assert rel_l2(out, x.grad) < 6e-3


@pytest.mark.parametrize("Mc,No,Ko,split", [(6336, 576, 192, 8), (6336, 192, 768, 16), (1000, 768, 192, 3),
                                            (50688, 576, 192, 29), (198, 136, 72, 1)])
def test_gemm_wgrad_mn_mn_splitk(Mc, No, Ko, split):
    """dW[No,Ko] = dY[Mc,No]^T @ X[Mc,Ko]: both operands MN-major, split-K with fp32 atomics."""
    dY = _rand(Mc, No, dtype=torch.bfloat16, seed=1)
    X = _rand(Mc, Ko, dtype=torch.bfloat16, seed=2)
    out = torch.zeros(No, Ko, dtype=torch.float32, device=DEV)
    ops.gemm(dY, X, No, Ko, Mc, a_mn=True, b_mn=True, out=out, split_k=split, epilogue=_lib.EPI_ATOMIC_ADD)
    ref = dY.float().t() @ X.float()
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 1e-4


def test_gemm_tokens_epilogue():
    B, P, T, prefix, D, K = 5, 196, 198, 2, 192, 768
    A = _rand(B * P, K, dtype=torch.bfloat16, seed=1)
    W = _rand(D, K, scale=0.05, dtype=torch.bfloat16, seed=2)
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
def test_layernorm_fwd_bwd(rows, dim):
    x = _rand(rows, dim, seed=1) * 2 + 0.5
    g = _rand(dim, seed=2) * 0.2 + 1
    b = _rand(dim, seed=3) * 0.1
    y, mean, rstd = ops.layernorm_fwd(x, g, b)
    ref = torch.nn.functional.layer_norm(x, (dim,), g, b, 1e-5)
    assert rel_l2(y, ref) < 4e-3
    twin = torch.empty(rows, dim, dtype=torch.bfloat16, device=DEV)
    yh, _, _ = ops.layernorm_fwd(x, g, b, dtype=torch.float16, y2=twin)
    assert rel_l2(yh, ref) < 5e-4 and torch.equal(twin, y)
    dy = _rand(rows, dim, dtype=torch.bfloat16, seed=4)
    dres = _rand(rows, dim, seed=5)
    dg = torch.zeros(dim, device=DEV); db = torch.zeros(dim, device=DEV); dc = torch.zeros(dim, device=DEV)
    dxb = torch.empty(rows, dim, dtype=torch.bfloat16, device=DEV)
    dx = ops.layernorm_bwd(dy, x, mean, rstd, g, dg, db, dres=dres, dx_bf16=dxb, dcolsum=dc)
    xr = x.clone().requires_grad_(True); gr = g.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (dim,), gr, br, 1e-5).backward(dy.float())
    ref_dx = xr.grad + dres
    torch.cuda.synchronize()
    assert rel_l2(dx, ref_dx) < 1e-5
    assert rel_l2(dxb, ref_dx) < 4e-3
    assert rel_l2(dg, gr.grad) < 1e-4
    assert rel_l2(db, br.grad) < 1e-4
    assert rel_l2(dc, ref_dx.sum(0)) < 1e-4


# ------------------------------------------------------------------ attention
def _attn_ref(qkv, B, N, H, scale):
    q, k, v = qkv.float().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).unbind(0)
    s = (q @ k.transpose(-2, -1)) * scale
    p = s.softmax(-1)
    o = (p @ v).transpose(1, 2).reshape(B, N, H * 64)
    return o, p, torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,N,H", [(2, 198, 3), (3, 197, 12), (1, 64, 1), (2, 577, 3), (1, 17, 2)])
def test_attention_fwd_bwd(B, N, H):
    scale = 64 ** -0.5
    qkv = _rand(B, N, 3 * H * 64, dtype=torch.bfloat16, seed=1)
    out, lse = ops.attention_fwd(qkv, B, N, H, scale)
    qr = qkv.float().requires_grad_(True)
    o_ref, p_ref, lse_ref = _attn_ref(qr, B, N, H, scale)
    torch.cuda.synchronize()
    assert rel_l2(out, o_ref) < 8e-3
    assert (lse - lse_ref).abs().max().item() < 2e-3
    outh = torch.empty(B, N, H * 64, dtype=torch.float16, device=DEV); twin = torch.empty_like(out)
    ops.attention_fwd(qkv, B, N, H, scale, out=outh, out2=twin)
    torch.cuda.synchronize()
    assert torch.equal(twin, out) and rel_l2(outh, o_ref) < 8e-3
    dout = _rand(B, N, H * 64, dtype=torch.bfloat16, seed=2)
    dqkv = ops.attention_bwd(qkv, out, dout, lse, B, N, H, scale)
    o_ref.backward(dout.float())
    torch.cuda.synchronize()
    assert rel_l2(dqkv, qr.grad) < 1.5e-2


def test_attention_probs_rows_sum_to_one():
    B, N, H = 2, 198, 3
    qkv = _rand(B, N, 3 * H * 64, dtype=torch.bfloat16, seed=1)
    probs = torch.empty(B, H, N, N, device=DEV)
    ops.attention_fwd(qkv, B, N, H, 0.125, probs=probs)
    _, p_ref, _ = _attn_ref(qkv, B, N, H, 0.125)
    torch.cuda.synchronize()
    assert (probs.sum(-1) - 1).abs().max().item() < 1e-5      # reference test_attention_quality.py:96-102
    assert probs.min().item() >= 0 and probs.max().item() <= 1
    assert (probs - p_ref).abs().max().item() < 1e-5


# ------------------------------------------------------------------ token plumbing
@pytest.mark.parametrize("B,C,S,P", [(3, 3, 224, 16), (2, 1, 256, 16), (2, 3, 64, 8), (1, 3, 64, 32)])
def test_patchify(B, C, S, P):
    img = _rand(B, C, S, S, seed=1)
    out = ops.patchify(img, P)
    ref = torch.nn.functional.unfold(img, P, stride=P).transpose(1, 2).reshape(-1, C * P * P)
    torch.cuda.synchronize()
    assert torch.equal(out, ref.to(torch.bfloat16))
    twin = torch.empty_like(out)
    outh = ops.patchify(img, P, out2=twin, dtype=torch.float16)
    torch.cuda.synchronize()
    assert torch.equal(outh, ref.to(torch.float16)) and torch.equal(twin, out)


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
    dpatch = torch.empty(B * (T - npre), D, dtype=torch.bfloat16, device=DEV)
    ops.tokens_bwd(dx, dpos, dcls, ddist, dpatch, dbias, npre)
    torch.cuda.synchronize()
    assert rel_l2(dpos, dx.sum(0)) < 1e-5
    assert rel_l2(dcls, dx[:, 0].sum(0)) < 1e-5 and rel_l2(ddist, dx[:, 1].sum(0)) < 1e-5
    assert rel_l2(dbias, dx[:, npre:].sum((0, 1))) < 1e-5
    assert torch.equal(dpatch, dx[:, npre:].reshape(-1, D).to(torch.bfloat16))


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
    dx = torch.full((B, T, D), 7.0, device=DEV); dxb = torch.empty(B, T, D, dtype=torch.bfloat16, device=DEV)
    z = lambda *s: torch.zeros(*s, device=DEV)
    dg, db_, dW0, db0, dcs = z(D), z(D), z(C, D), z(C), z(D)
    dW1, db1 = (z(C, D), z(C)) if n_heads == 2 else (None, None)
    ops.head_bwd(dl0, dl1, xhat, rstd, g, b, W0, W1, dx, dxb, dg, db_, dW0, db0, dW1, db1, dcs, T, n_heads)
    torch.cuda.synchronize()
    assert rel_l2(dx, xr.grad) < 1e-5 and dx[:, n_heads:].abs().max().item() == 0
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


# ------------------------------------------------------------------ AdamW + clip
@pytest.mark.parametrize("max_norm", [0.0, 1.0])
def test_adamw_matches_torch(max_norm):
    sizes = [(192, 768), (192,), (2, 192), (1, 198, 192), (5,)]
    lr, wd_list, scale_list = 1e-3, [0.05, 0.0, 0.05, 0.0, 0.05], [1.0, 0.75, 0.5, 1.0, 0.1]
    offs, total = [], 0
    for s in sizes:
        offs.append(total); total += (math.prod(s) + 127) // 128 * 128
    flat_p = torch.zeros(total, device=DEV); flat_g = torch.zeros(total, device=DEV)
    m = torch.zeros(total, device=DEV); v = torch.zeros(total, device=DEV)
    p16 = torch.zeros(total, dtype=torch.bfloat16, device=DEV)
    ph16 = torch.zeros(total, dtype=torch.float16, device=DEV)
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
        ops.adamw_step(flat_p, flat_g, m, v, p16, ph16, chunk_off, chunk_len, c_scale, c_wd, state, 0.9, 0.999, 1e-8, max_norm)
    torch.cuda.synchronize()
    for r, s, o in zip(refs, sizes, offs):
        n = math.prod(s)
        assert (flat_p[o:o + n] - r.detach().flatten()).abs().max().item() < 2e-6
        assert torch.equal(p16[o:o + n], flat_p[o:o + n].to(torch.bfloat16))
        assert torch.equal(ph16[o:o + n], flat_p[o:o + n].to(torch.float16))
    assert state[0].item() == 3.0


# ------------------------------------------------------------------ helpers
def test_cast_and_colsum():
    x = _rand(1000, 576, seed=1)
    xb = ops.cast_bf16(x)
    assert torch.equal(xb, x.to(torch.bfloat16))
    assert torch.equal(ops.cast_fp16(x), x.to(torch.float16))
    out = torch.zeros(576, device=DEV)
    ops.colsum_bf16(xb, out)
    torch.cuda.synchronize()
    assert rel_l2(out, xb.float().sum(0)) < 1e-5


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


# ------------------------------------------------------------------ mixed 16-bit operand formats
@pytest.mark.parametrize("adt,bdt", [(torch.float16, torch.float16), (torch.bfloat16, torch.bfloat16)])
def test_gemm_operand_formats(adt, bdt):
    """kind::f16 takes fp16 x fp16 or bf16 x bf16.  (Mixing the two in ONE tcgen05.mma was probed on B200 and raises
    cudaErrorIllegalInstruction, so the engine keeps a bf16 twin of every forward operand for backward.)"""
    M, N, K = 1000, 192, 320
    A32, B32 = _rand(M, K, seed=1), _rand(N, K, scale=0.05, seed=2)
    A, B = A32.to(adt), B32.to(bdt)
    out = torch.empty(M, N, dtype=torch.float32, device=DEV)
    ops.gemm(A, B, M, N, K, out=out)
    torch.cuda.synchronize()
    assert rel_l2(out, A.float() @ B.float().t()) < 1e-5
    # dgrad layout (B MN-major) and wgrad layout (both MN-major, split-K) with mixed formats
    Bt = _rand(K, N, scale=0.05, seed=3).to(bdt)
    out16 = torch.empty(M, N, dtype=torch.float16, device=DEV)
    ops.gemm(A, Bt, M, N, K, b_mn=True, out=out16)
    torch.cuda.synchronize()
    assert rel_l2(out16, A.float() @ Bt.float()) < 1e-3
    dY, X = _rand(M, N, seed=4).to(adt), _rand(M, K, seed=5).to(bdt)
    dW = torch.zeros(N, K, device=DEV)
    ops.gemm(dY, X, N, K, M, a_mn=True, b_mn=True, out=dW, split_k=4, epilogue=_lib.EPI_ATOMIC_ADD)
    torch.cuda.synchronize()
    assert rel_l2(dW, dY.float().t() @ X.float()) < 1e-4
