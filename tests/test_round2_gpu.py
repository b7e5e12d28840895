"""Round-2 GPU parity tests: the gray-tile input path, ViT-B/16 at larger batches, loss-scale overflow inside a captured
graph, optimizer checkpointing, the assembled fold-sharded ensemble (BASELINE config 5), data-parallel gradient parity on
two ranks, and the reference's own structural tests ported onto the drop-in classes.

Tolerances are BASELINE.json's: logits <= 2e-3 max-abs, per-parameter gradients <= 1e-2 relative L2, identical top-1."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import thyroid_vit_cnn_comparison_b200 as tv  # noqa: E402,F401
from thyroid_vit_cnn_comparison_b200 import vit as V, training as TR, optim as OPT, ops, ensemble as ENS, parallel  # noqa: E402
from thyroid_vit_cnn_comparison_b200.engine import GraySpec  # noqa: E402
from oracle import vit_oracle as O  # noqa: E402
from test_model_gpu import build, run_gpu, rel_l2, LOGIT_TOL, GRAD_TOL  # noqa: E402

DEV = "cuda"


# ------------------------------------------------------------------ gray tiles -> patch matrix
@pytest.mark.parametrize("kind", ["u16", "f16", "f32"])
@pytest.mark.parametrize("norm", [False, True])
def test_tiles_to_patches_equals_finish_tiles_plus_patchify(kind, norm):
    """vitk_tiles_to_patches == vitk_finish_tiles (replicate + Normalize [+ percentile clamp]) followed by vitk_patchify,
    bit for bit, for raw uint16 tiles (/65535 on the device), fp16 and fp32 tiles."""
    g = torch.Generator().manual_seed(5)
    B, S, P, C = 5, 64, 16, 3
    if kind == "u16":
        tiles = torch.randint(0, 65536, (B, S, S), generator=g, dtype=torch.int32).to(torch.uint16).to(DEV)
        gray = ops.resize_u16(tiles, S, S)                                   # same size: only the / 65535 (dataset.py:549)
        assert torch.equal(gray.cpu(), tiles.cpu().to(torch.float32) / 65535.0)
    elif kind == "f16":
        tiles = torch.rand(B, S, S, generator=g).to(torch.float16).to(DEV)
        gray = tiles.float()
    else:
        tiles = torch.rand(B, S, S, generator=g).to(DEV)
        gray = tiles
    mean, std = ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)) if norm else (None, None)
    bounds = ops.percentile_bounds(gray, 0.01, 0.99) if norm else None
    ref = ops.patchify(ops.finish_tiles(gray, C, bounds=bounds, mean=mean, std=std), P)
    got = ops.tiles_to_patches(tiles, C, P, bounds=bounds, mean=mean, std=std)
    torch.cuda.synchronize()
    assert got.shape == ref.shape == (B * 16, C * P * P)
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))
    got4 = ops.tiles_to_patches(tiles.unsqueeze(1), C, P, bounds=bounds, mean=mean, std=std)     # [B,1,H,W] accepted
    assert torch.equal(got4.view(torch.int16), ref.view(torch.int16))
    with pytest.raises(RuntimeError):
        ops.tiles_to_patches(tiles, C, 12)                                   # P % 8 != 0


@pytest.mark.parametrize("use_graph", [False, True])
def test_train_step_from_gray_tiles_equals_the_replicated_batch(use_graph):
    """TrainStep(input_format='gray') on raw uint16 tiles == TrainStep on the loader's fp32 [B,3,H,W] batch (tile/65535
    replicated to 3 channels): bit-identical step statistics (the forward is deterministic); gradients / parameters equal up to
    the summation order of the split-K / atomic weight-gradient accumulation (1e-5 relative)."""
    cfg = O.VitConfig(img_size=64, embed_dim=128, depth=2, num_heads=2)
    g = torch.Generator().manual_seed(9)
    tiles = torch.randint(0, 65536, (8, 64, 64), generator=g, dtype=torch.int32).to(torch.uint16)
    y = torch.randint(0, 2, (8,), generator=g)
    x = (tiles.to(torch.float32) / 65535.0).unsqueeze(1).expand(-1, 3, -1, -1).contiguous()
    results = []
    for fmt in ("gray", "nchw"):
        model, _ = build(cfg, 7)
        model.train()
        opt = OPT.FusedAdamW(model, lr=1e-3, weight_decay=0.05, max_grad_norm=1.0)
        step = TR.TrainStep(model, opt, 8, mode="ce", use_graph=use_graph, input_format=fmt)
        first = step((tiles if fmt == "gray" else x).pin_memory(), y.pin_memory()).cpu().clone()
        stats = step((tiles if fmt == "gray" else x).pin_memory(), y.pin_memory()).cpu()
        torch.cuda.synchronize()
        eng = model._engine
        results.append((first, stats.clone(), eng.flat.grads.clone(), eng.flat.params.clone()))
    (f0, s0, g0, p0), (f1, s1, g1, p1) = results
    assert torch.equal(f0, f1)                                   # the forward is deterministic: same patch matrix, same loss
    # step 2 starts from parameters equal up to the summation order of step 1's atomically accumulated gradients; a last-bit
    # difference in a master weight can flip the rounding of its fp16 shadow, so step 2 agrees to ~1e-3 of the gradient norm
    assert (s0 - s1).abs().max().item() < 1e-4
    assert rel_l2(g0, g1) < 2e-3 and (p0 - p1).abs().max().item() < 1e-5
    assert torch.isfinite(s0).all() and g0.abs().sum().item() > 0


# ------------------------------------------------------------------ ViT-B/16 at larger batches
def test_vit_base_b16_vs_oracle():
    """ViT-B/16 (BASELINE config 4 model) logits + every parameter gradient against the CPU oracle at batch 16."""
    cfg = O.VIT_BASE
    model, sd = build(cfg, 42)
    x, y = O.seeded_batch(cfg, 16, 42)
    loss, outs, grads = run_gpu(model, x, y)
    ref_loss, ref_out, ref_grads = O.train_step(sd, x, y, cfg)
    assert (outs[0] - ref_out.detach()).abs().max().item() < LOGIT_TOL
    assert torch.equal(outs[0].argmax(1), ref_out.detach().argmax(1))
    assert abs(loss - ref_loss.item()) < 2e-3
    worst = max((rel_l2(grads[n], g), n) for n, g in ref_grads.items() if g is not None)
    assert worst[0] < GRAD_TOL, worst


def test_vit_base_full_batch_256_properties():
    """ViT-B/16 at the bench size (batch 256): batch independence (bit-identical logits for a 32-image slice run alone),
    determinism, and linearity of the mean-reduced gradient over 8 sub-batches (D = 768 exercises the cta_group::2 GEMMs)."""
    cfg = O.VIT_BASE
    model, _ = build(cfg, 42)
    x, y = O.seeded_batch(cfg, 256, 43)
    loss, outs, grads = run_gpu(model, x, y)
    _, outs2, _ = run_gpu(model, x, y)
    assert torch.equal(outs[0], outs2[0])
    _, outs32, _ = run_gpu(model, x[64:96], y[64:96])
    assert torch.equal(outs[0][64:96], outs32[0])
    acc = {n: torch.zeros_like(g) for n, g in grads.items() if g is not None}
    for k in range(8):
        _, _, g = run_gpu(model, x[32 * k:32 * k + 32], y[32 * k:32 * k + 32])
        for n in acc:
            acc[n] += g[n] / 8
    worst = max((rel_l2(grads[n], acc[n]), n) for n in acc)
    assert worst[0] < 2e-3, worst
    assert torch.isfinite(outs[0]).all() and abs(loss - 0.6931) < 0.2


# ------------------------------------------------------------------ loss-scale overflow inside the captured graph
def test_forced_fp16_overflow_inside_captured_graph_skips_the_step():
    """A loss scale far too large makes the fp16 activation gradients overflow: the captured step must leave parameters, both
    Adam moments and the step counter bit-identical, halve S and count the skip; the following steps train normally."""
    cfg = O.VitConfig(img_size=64, embed_dim=128, depth=2, num_heads=2)
    model, _ = build(cfg, 7)
    model.train()
    opt = OPT.FusedAdamW(model, lr=1e-3, weight_decay=0.05, max_grad_norm=1.0)
    step = TR.TrainStep(model, opt, 8, mode="ce", use_graph=True)
    x, y = O.seeded_batch(cfg, 8, 3)
    for _ in range(2):                                            # capture both input slots on clean steps
        step(x.cuda(), y.cuda())
    torch.cuda.synchronize()
    eng = model._engine
    assert step.graphs[0] is not None and step.graphs[1] is not None and eng.amp[3].item() == 0.0
    p0, m0, v0 = eng.flat.params.clone(), opt.exp_avg.clone(), opt.exp_avg_sq.clone()
    w0, t0 = eng.flat.w16.clone(), opt.dev_state[0].item()
    big = 2.0 ** 40                                               # S * dlogits >> 65504
    eng.amp[0], eng.amp[1] = big, 1.0 / big
    stats = step(x.cuda(), y.cuda()).cpu()
    torch.cuda.synchronize()
    assert torch.isfinite(stats[0])                               # the forward / loss are unaffected by the loss scale
    assert torch.equal(eng.flat.params, p0) and torch.equal(opt.exp_avg, m0) and torch.equal(opt.exp_avg_sq, v0)
    assert torch.equal(eng.flat.w16, w0) and opt.dev_state[0].item() == t0
    assert eng.amp[0].item() == big / 2 and eng.amp[3].item() == 1.0 and eng.amp[4].item() == 1.0 and eng.amp[2].item() == 0.0
    eng.amp[0], eng.amp[1] = 65536.0, 1.0 / 65536.0
    step(x.cuda(), y.cuda())
    torch.cuda.synchronize()
    assert opt.dev_state[0].item() == t0 + 1 and not torch.equal(eng.flat.params, p0) and eng.amp[4].item() == 0.0


def test_lightning_path_ticks_the_loss_scale_once_per_optimizer_step():
    """model(x) -> loss.backward() -> FusedAdamW.step(): exactly ONE loss-scale bookkeeping update per step (the optimizer's),
    and an overflowed step is skipped by the optimizer (no weight decay, no moment decay, no step increment)."""
    cfg = O.VitConfig(img_size=64, embed_dim=64, depth=1, num_heads=1)
    model, _ = build(cfg, 3)
    model.train()
    opt = OPT.FusedAdamW(model, lr=1e-3, weight_decay=0.1, max_grad_norm=1.0)
    eng = model._engine
    x, y = O.seeded_batch(cfg, 4, 1)
    for i in range(3):
        opt.zero_grad()
        loss, _ = TR.fused_cross_entropy(model(x.cuda()), y.cuda())
        loss.backward()
        opt.step()
        torch.cuda.synchronize()
        assert eng.amp[2].item() == float(i + 1) and opt.dev_state[0].item() == float(i + 1)
    p0, m0 = eng.flat.params.clone(), opt.exp_avg.clone()
    eng.amp[0], eng.amp[1] = 2.0 ** 40, 2.0 ** -40
    opt.zero_grad()
    loss, _ = TR.fused_cross_entropy(model(x.cuda()), y.cuda())
    loss.backward()
    opt.step()
    torch.cuda.synchronize()
    assert torch.equal(eng.flat.params, p0) and torch.equal(opt.exp_avg, m0) and opt.dev_state[0].item() == 3.0
    assert eng.amp[3].item() == 1.0 and eng.amp[0].item() == 2.0 ** 39


# ------------------------------------------------------------------ optimizer checkpointing
def test_fused_adamw_state_dict_roundtrip_and_torch_adamw_interop():
    cfg = O.VitConfig(img_size=64, embed_dim=64, depth=2, num_heads=1)
    x, y = O.seeded_batch(cfg, 4, 2)

    def make():
        model, _ = build(cfg, 5)
        model.train()
        opt = OPT.FusedAdamW(model, lr=1e-3, weight_decay=0.05, max_grad_norm=1.0)
        return model, opt, TR.TrainStep(model, opt, 4, mode="ce")

    ma, oa, sa = make()
    for _ in range(3):
        sa(x.cuda(), y.cuda())
    torch.cuda.synchronize()
    ckpt = {"model": {k: v.clone() for k, v in ma.state_dict().items()}, "opt": oa.state_dict()}
    assert set(ckpt["opt"]) >= {"state", "param_groups", "vitk"} and len(ckpt["opt"]["state"]) > 0
    st0 = ckpt["opt"]["state"][0]
    assert set(st0) == {"step", "exp_avg", "exp_avg_sq"} and float(st0["step"]) == 3.0       # torch.optim.AdamW's layout
    import io
    buf = io.BytesIO()
    torch.save(ckpt, buf)                                           # what a Lightning checkpoint does with it
    buf.seek(0)
    ckpt = torch.load(buf, weights_only=False)
    mb, ob, sb = make()
    mb.load_state_dict(ckpt["model"])
    ob.load_state_dict(ckpt["opt"])
    assert torch.equal(ob.exp_avg, oa.exp_avg) and torch.equal(ob.exp_avg_sq, oa.exp_avg_sq)
    assert torch.equal(ob.dev_state, oa.dev_state) and torch.equal(mb._engine.amp, ma._engine.amp)
    sa(x.cuda(), y.cuda())
    sb(x.cuda(), y.cuda())
    torch.cuda.synchronize()
    # resumed run == uninterrupted run, up to the summation order of the atomically accumulated weight gradients: where a
    # gradient element is itself rounding noise (|g| ~ eps) Adam's normalised update moves by a fraction of lr = 1e-3
    assert (ma._engine.flat.params - mb._engine.flat.params).abs().max().item() < 2.5e-4
    assert (ma._engine.flat.params - mb._engine.flat.params).abs().mean().item() < 1e-6
    assert rel_l2(oa.exp_avg, ob.exp_avg) < 1e-5 and oa.dev_state[0].item() == 4.0 == ob.dev_state[0].item()
    # a checkpoint written by torch.optim.AdamW (what the reference's Lightning run saves) resumes here
    mc, oc, _ = make()
    ref_opt = torch.optim.AdamW([p for p in mc.parameters()], lr=1e-3, weight_decay=0.05)
    mc.zero_grad()
    loss, _ = TR.fused_cross_entropy(mc(x.cuda()), y.cuda())
    loss.backward()
    ref_opt.step()
    torch.cuda.synchronize()
    oc.load_state_dict(ref_opt.state_dict())
    n = "blocks.1.mlp.fc1.weight"
    p = dict(mc.named_parameters())[n]
    assert torch.equal(mc._engine.flat.view(oc.exp_avg, n), ref_opt.state[p]["exp_avg"]) and oc.dev_state[0].item() == 1.0


def test_loss_kernel_counts_out_of_range_labels_and_poisons_the_step():
    logits = torch.randn(6, 2, device=DEV)
    labels = torch.tensor([0, 1, 2, 1, -1, 0], device=DEV)
    out, d0, _ = ops.loss_fwd_bwd(logits, None, None, labels, mode=0, w_cls=1.0, w_dist=0.0)
    torch.cuda.synchronize()
    assert out[6].item() == 2.0 and torch.isnan(out[0])
    assert torch.isnan(d0[2]).all() and torch.isnan(d0[4]).all() and torch.isfinite(d0[[0, 1, 3, 5]]).all()
    out, d0, _ = ops.loss_fwd_bwd(logits, None, None, labels.clamp(0, 1), mode=0, w_cls=1.0, w_dist=0.0)
    assert out[6].item() == 0.0 and torch.isfinite(out[0]) and torch.isfinite(d0).all()
    ref = torch.nn.functional.cross_entropy(logits, labels.clamp(0, 1))
    assert abs(out[0].item() - ref.item()) < 1e-6


# ------------------------------------------------------------------ BASELINE config 5: fold-sharded ensemble + rollout
def _ensemble_reference(cfg, seeds, x, weights):
    logits, grids = [], []
    with torch.no_grad():
        for s in seeds:
            sd = O.seeded_state_dict(cfg, s)
            maps = []
            logits.append(O.forward(sd, x, cfg, training=False, attn_out=maps))
            r = O.attention_rollout(torch.stack(maps), "mean")
            grids.append(r[:, 0, cfg.num_prefix:])                         # class-token row over the patch columns
    probs, preds = O.ensemble_predict(torch.stack(logits), weights)
    g = int(round(cfg.num_patches ** 0.5))
    return torch.stack(logits), probs, preds, torch.stack(grids).reshape(len(seeds), x.shape[0], g, g)


def test_ensemble_inference_matches_oracle_deit_tiny_five_folds():
    """F = 5 DeiT-tiny members (seeds 42+f), uniform weights: bit-identical top-1 against the oracle's evaluate_ensemble
    restatement, probabilities within the logit tolerance, per-fold rollout grids against the spec restatement; a
    non-uniform weight vector (the reference script's 0.5/0.25/0.25 pattern) as well."""
    cfg = O.DEIT_TINY
    seeds = [42 + f for f in range(5)]
    x, _ = O.seeded_batch(cfg, 16, 42)
    members = []
    for s in seeds:
        m, _ = build(cfg, s)
        members.append(m)
    ens = ENS.EnsembleInference(members, rollout=True)
    out = ens(x.cuda())
    torch.cuda.synchronize()
    w = torch.full((5,), 0.2)
    ref_logits, ref_probs, ref_preds, ref_grids = _ensemble_reference(cfg, seeds, x, w)
    assert out["logits"].shape == (5, 16, 2) and out["rollout"].shape == (5, 16, 14, 14)
    assert (out["logits"].cpu() - ref_logits).abs().max().item() < LOGIT_TOL
    assert (out["probs"].cpu() - ref_probs).abs().max().item() < 1e-3
    assert torch.equal(out["preds"].cpu(), ref_preds)                                    # bit-identical top-1
    assert (out["rollout"].cpu() - ref_grids).abs().max().item() < 2e-4
    assert (out["rollout"].sum((-1, -2)).cpu() <= 1.0 + 1e-4).all()                        # a sub-row of a row-stochastic matrix
    w2 = [0.4, 0.15, 0.15, 0.15, 0.15]
    out2 = ENS.EnsembleInference(members, weights=w2)(x.cuda())
    _, p2, y2, _ = _ensemble_reference(cfg, seeds, x, torch.tensor(w2))
    assert torch.equal(out2["preds"].cpu(), y2) and (out2["probs"].cpu() - p2).abs().max().item() < 1e-3
    # gray-tile input gives the same ensemble as the replicated batch
    tiles = torch.randint(0, 65536, (4, 224, 224), dtype=torch.int32).to(torch.uint16).cuda()
    x3 = (tiles.float() / 65535.0).unsqueeze(1).expand(-1, 3, -1, -1).contiguous()
    a, b = ens(tiles, gray=GraySpec()), ens(x3)
    assert torch.equal(a["logits"], b["logits"]) and torch.equal(a["preds"], b["preds"])
    with pytest.raises(AssertionError):
        ENS.EnsembleInference(members, weights=[0.5, 0.5])
    # EnsembleTeacher (src/utils/models.py:231-283): normalised weights, weighted LOGIT average
    t = ENS.EnsembleTeacher(members[:3], weights=[2.0, 1.0, 1.0])
    tl = t(x[:4].cuda())
    ref_t = O.ensemble_teacher_logits(ref_logits[:3, :4], [2.0, 1.0, 1.0])
    assert (tl.cpu() - ref_t).abs().max().item() < LOGIT_TOL
    # create_attention_rollout on the maps the model API returns
    members[0].eval()
    with torch.no_grad():
        members[0](x[:2].cuda())
    maps = members[0].get_attention_maps()
    assert maps.shape == (12, 2, 3, 198, 198)
    r = ENS.create_attention_rollout(maps.cuda(), "max")
    assert (r.cpu() - O.attention_rollout(maps, "max")).abs().max().item() < 1e-5
    with pytest.raises(ValueError):
        ENS.create_attention_rollout(maps.cuda(), "median")


# ------------------------------------------------------------------ two ranks (needs 2 GPUs)
def _two_rank_worker(rank, world, port, tmp, graph):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cfg = O.VitConfig(img_size=64, embed_dim=128, depth=3, num_heads=2)
        x, y = O.seeded_batch(cfg, 16, 77)
        # ---- data parallel: each rank steps on its half, gradients all-reduced in >= 3 overlapped buckets
        model, _ = build(cfg, 7)
        model.train()
        opt = OPT.FusedAdamW(model, lr=1e-3, weight_decay=0.05, max_grad_norm=1.0)
        red = parallel.BucketedAllReduce(min_buckets=4)
        step = TR.TrainStep(model, opt, 8, mode="ce", reducer=red, use_graph=graph)
        lo = 8 * rank
        for _ in range(2):
            stats = step(x[lo:lo + 8].cuda(), y[lo:lo + 8].cuda()).clone()
        torch.cuda.synchronize()
        res = {"grads": model._engine.flat.grads.cpu(), "params": model._engine.flat.params.cpu(), "buckets": len(red.buckets),
               "early": sum(1 for s, _ in red.launched_log if s != "embed")}
        # ---- fold-sharded ensemble: 5 folds over 2 ranks
        cfg_e = O.VitConfig(img_size=64, embed_dim=64, depth=2, num_heads=1)
        mine = parallel.shard_folds(5, rank, world)
        members = [build(cfg_e, 42 + f)[0] for f in mine]
        xe, _ = O.seeded_batch(cfg_e, 6, 9)
        out = ENS.EnsembleInference(members, num_folds=5, rollout=True)(xe.cuda())
        torch.cuda.synchronize()
        res.update(ens_logits=out["logits"].cpu(), ens_preds=out["preds"].cpu(), ens_rollout=out["rollout"].cpu(), folds=mine)
        torch.save(res, Path(tmp) / f"r{rank}.pt")
        # teardown as in bench.py: captured graphs pin NCCL work, so they go first; a watchdog ends the process should the
        # communicator teardown itself block (the results are already on disk)
        import threading
        import time

        def _bail():
            time.sleep(15)
            os._exit(0)
        threading.Thread(target=_bail, daemon=True).start()
        step.graphs = [None, None]
        del step
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("graph", [False, True])
def test_two_rank_data_parallel_gradients_and_fold_sharded_ensemble(tmp_path, graph):
    """2 ranks x 8 images must give the single-rank gradients (and updated parameters) of the 16 concatenated images; the
    fold-sharded ensemble must equal the single-process ensemble on every rank."""
    import torch.multiprocessing as mp
    port = 29700 + (os.getpid() % 1000) + (1 if graph else 0)
    mp.spawn(_two_rank_worker, args=(2, port, str(tmp_path), graph), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f"r{r}.pt", weights_only=False) for r in range(2))
    assert torch.equal(r0["grads"], r1["grads"]) and torch.equal(r0["params"], r1["params"])      # replicas stay in lockstep
    assert r0["buckets"] >= 3 and r0["early"] >= 2                                                 # overlap: buckets left before 'embed'
    cfg = O.VitConfig(img_size=64, embed_dim=128, depth=3, num_heads=2)
    x, y = O.seeded_batch(cfg, 16, 77)
    model, _ = build(cfg, 7)
    model.train()
    opt = OPT.FusedAdamW(model, lr=1e-3, weight_decay=0.05, max_grad_norm=1.0)
    step = TR.TrainStep(model, opt, 16, mode="ce")
    for _ in range(2):
        step(x.cuda(), y.cuda())
    torch.cuda.synchronize()
    eng = model._engine
    for n in eng.flat.order:
        off, shape = eng.flat.offsets[n]
        a = r0["grads"][off:off + shape.numel()]
        b = eng.g(n).reshape(-1).cpu()
        assert rel_l2(a, b) < 2e-3, (n, rel_l2(a, b))
    assert (r0["params"] - eng.flat.params.cpu()).abs().max().item() < 2.5e-3                       # two Adam steps of lr 1e-3
    cfg_e = O.VitConfig(img_size=64, embed_dim=64, depth=2, num_heads=1)
    members = [build(cfg_e, 42 + f)[0] for f in range(5)]
    xe, _ = O.seeded_batch(cfg_e, 6, 9)
    single = ENS.EnsembleInference(members, rollout=True)(xe.cuda())
    for r in (r0, r1):
        assert torch.equal(r["ens_logits"], single["logits"].cpu()) and torch.equal(r["ens_preds"], single["preds"].cpu())
        assert torch.equal(r["ens_rollout"], single["rollout"].cpu())
    assert r0["folds"] == [0, 2, 4] and r1["folds"] == [1, 3]


# ------------------------------------------------------------------ the reference's structural tests, on the drop-in classes
def _simple_vit(**kwargs):
    """tests/test_vision_transformer_base.py:173-208: a VisionTransformerBase subclass that builds its blocks from hparams."""
    class SimpleViT(V.VisionTransformerBase):
        def __init__(self, **kw):
            super().__init__(**kw)
            dpr = [v.item() for v in torch.linspace(0, self.hparams.drop_path_rate, self.hparams.depth)]
            self.blocks = nn.Sequential(*[
                V.Block(dim=self.embed_dim, num_heads=self.hparams.num_heads, mlp_ratio=self.hparams.mlp_ratio,
                        qkv_bias=self.hparams.qkv_bias, drop=self.hparams.drop_rate, attn_drop=self.hparams.attn_drop_rate,
                        drop_path=dpr[i], store_attention=self.hparams.store_attention)
                for i in range(self.hparams.depth)])
    kw = dict(img_size=224, patch_size=16, in_chans=1, num_classes=2, embed_dim=192, depth=2, num_heads=3, mlp_ratio=4.0,
              qkv_bias=True, drop_path_rate=0.1)
    kw.update(kwargs)
    return SimpleViT(**kw).cuda()


def _synthetic_batch(b=4, s=224):
    g = torch.Generator().manual_seed(0)
    return torch.randn(b, 1, s, s, generator=g).cuda(), torch.randint(0, 2, (b,), generator=g).cuda()


def test_reference_structural_suite_on_the_drop_in_base_class():
    images, labels = _synthetic_batch()
    model = _simple_vit()
    assert model(images).shape == (4, 2)                                                 # test_forward_pass (:210-216)
    model.eval()
    with torch.no_grad():
        assert model.extract_features(images).shape == (4, 192)                          # test_feature_extraction (:218-224)
    model.train()
    loss = model.training_step((images, labels), 0)                                      # test_training_step (:226-232)
    assert isinstance(loss, torch.Tensor) and loss.ndim == 0
    res = model.validation_step((images, labels), 0)                                     # test_validation_step (:234-240)
    assert "val_loss" in res and "val_acc" in res
    assert isinstance(_simple_vit(pos_embed_type="learnable").pos_embed, nn.Parameter)   # test_position_embeddings (:242-252)
    assert not isinstance(_simple_vit(pos_embed_type="sinusoidal").pos_embed, nn.Parameter)
    assert _simple_vit(class_token=True, pool_type="cls")(images).shape == (4, 2)        # test_pooling_strategies (:254-266)
    assert _simple_vit(class_token=False, pool_type="gap")(images).shape == (4, 2)
    m = _simple_vit(store_attention=True).eval()                                         # test_attention_visualization (:268-282)
    with torch.no_grad():
        m(images)
        maps = m.get_attention_maps()
    assert maps is not None and maps.shape[0] == 2 and maps.shape[1] == 4
    groups = _simple_vit(depth=4).get_parameter_groups(weight_decay=0.05, layer_decay=0.75)   # test_parameter_groups (:284-305)
    assert len(groups) > 0 and any("blocks" in g["name"] and "lr_scale" in g for g in groups)


@pytest.mark.parametrize("img_size", [224, 256, 384])
def test_reference_different_image_sizes(img_size):                                      # :311-321
    model = _simple_vit(img_size=img_size, patch_size=16)
    x = torch.randn(2, 1, img_size, img_size).cuda()
    assert model(x).shape == (2, 2)


def test_reference_gradient_flow():                                                      # :323-340
    images, labels = _synthetic_batch()
    model = _simple_vit()
    loss = torch.nn.functional.cross_entropy(model(images), labels)
    loss.backward()
    torch.cuda.synchronize()
    for name, p in model.named_parameters():
        if p.requires_grad and "quality_score" not in name:
            assert p.grad is not None, f"No gradient for {name}"
            assert not torch.isnan(p.grad).any(), f"NaN gradient for {name}"
            assert p.grad.abs().sum().item() > 0, f"Zero gradient for {name}"


@pytest.mark.parametrize("batch_size", [1, 2, 8, 16])
def test_reference_batch_sizes(batch_size):                                              # :342-349
    model = _simple_vit()
    assert model(torch.randn(batch_size, 1, 224, 224).cuda()).shape == (batch_size, 2)


def test_drop_path_and_patch_embed_stand_alone():
    """tests/test_vision_transformer_base.py:150-168 (DropPath) and :60-84 (PatchEmbed returns (tokens, quality scores))."""
    x = torch.randn(4, 16, 32, device=DEV)
    dp = V.DropPath(0.5)
    dp.eval()
    assert torch.allclose(dp(x), x)
    dp.train()
    torch.manual_seed(0)
    out = dp(x)
    assert out.shape == x.shape
    per_sample = out.reshape(4, -1).abs().sum(1) / x.reshape(4, -1).abs().sum(1)
    assert all(abs(v) < 1e-6 or abs(v - 2.0) < 1e-5 for v in per_sample.tolist())        # dropped, or scaled by 1 / keep
    pe = V.PatchEmbed(img_size=64, patch_size=16, in_chans=1, embed_dim=64).cuda()
    tok, q = pe(torch.rand(2, 1, 64, 64, device=DEV))
    assert tok.shape == (2, 16, 64) and q.shape == (2, 16) and (q >= 0).all() and (q <= 1).all()
    tok2, q2 = V.PatchEmbed(img_size=64, patch_size=16, in_chans=1, embed_dim=64, quality_aware=False).cuda()(torch.rand(2, 1, 64, 64, device=DEV))
    assert q2 is None


@pytest.mark.parametrize("variant", ["cls", "gap_rep"])
def test_forward_features_trains_with_autograd(variant):
    """vision_transformer_base.py:440-479 in training mode: the pre-head feature carries gradients into the encoder (a custom
    head on top of forward_features); every encoder gradient against the oracle's autograd through the same feature."""
    kw = dict(pool_type="gap", representation_size=64) if variant == "gap_rep" else {}
    cfg = O.VitConfig(img_size=64, embed_dim=64, depth=2, num_heads=1, is_deit=False, distilled=False,
                      pool_type=kw.get("pool_type", "cls"), representation_size=kw.get("representation_size", 0))
    model, sd = build(cfg, 13)
    model.train()
    x, _ = O.seeded_batch(cfg, 6, 13)
    # a feature gradient of the size a mean-reduced loss produces (an O(1) gradient times the initial loss scale 65536 would
    # overflow fp16 and -- GradScaler semantics -- zero this step's gradients while the scale halves)
    proj = torch.randn(64, generator=torch.Generator().manual_seed(1)) * 1e-3
    model.zero_grad()
    feats, q = model.forward_features(x.cuda())
    assert q is None and feats.shape == (6, 64) and feats.requires_grad
    (feats * proj.cuda()).sum().backward()
    torch.cuda.synchronize()
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = O.pooled_features(leaves, O.forward_tokens(leaves, x, cfg), cfg)
    (ref * proj).sum().backward()
    assert (feats.detach().cpu() - ref.detach()).abs().max().item() < LOGIT_TOL
    for n, p in model.named_parameters():
        g = leaves[n].grad
        if g is None:
            assert p.grad is None or n.startswith("head") or p.grad.abs().sum().item() == 0, n    # the head is not on this path
            continue
        assert rel_l2(p.grad, g) < GRAD_TOL, (n, rel_l2(p.grad, g))


# ------------------------------------------------------------------ last block on the classifier's rows
@pytest.mark.gpu
@pytest.mark.parametrize("B,T,n,dt", [(5, 198, 2, torch.float16), (3, 197, 1, torch.float32), (256, 198, 2, torch.float32)])
def test_gather_and_expand_rows(B, T, n, dt):
    """vitk_gather_rows / vitk_expand_rows: leading token rows of every image, compact <-> dense (bit-exact copies)."""
    D = 192
    x = torch.randn(B, T, D, device="cuda").to(dt)
    c = torch.empty(B, n, D, dtype=dt, device="cuda")
    ops.gather_rows(x, n, c)
    assert torch.equal(c, x[:, :n])
    dense = torch.full((B, T, D), 7.0, dtype=dt, device="cuda")
    ops.expand_rows(c, n, dense)
    assert torch.equal(dense[:, :n], x[:, :n]) and dense[:, n:].abs().max().item() == 0.0
    with pytest.raises(ValueError):
        ops.gather_rows(x, n, torch.empty(B, n + 1, D, dtype=dt, device="cuda"))


@pytest.mark.gpu
@pytest.mark.parametrize("deit", [True, False])
def test_last_block_on_classifier_rows_equals_dense(deit):
    """engine.cls_rows_last_block: the last block's attn.proj / norm2 / Mlp on tokens 0..n_out-1 only gives the logits and
    every parameter gradient of the dense computation (vision_transformer_base.py:274-285,474-479; deit_models.py:224-235)."""
    from thyroid_vit_cnn_comparison_b200 import vit
    kw = dict(img_size=64, patch_size=16, in_chans=3, num_classes=6, embed_dim=128, depth=3, num_heads=2, mlp_ratio=4.0)
    outs = {}
    for mode in ("rows", "dense"):
        torch.manual_seed(11)
        m = (vit.DeiT(distilled=True, **kw) if deit else vit.VisionTransformer(**kw)).cuda().train()
        eng = m._ensure_engine()
        eng.cls_rows_last_block = mode == "rows"
        x = torch.randn(7, 3, 64, 64, generator=torch.Generator().manual_seed(5)).cuda()
        out = m(x)
        logits = torch.cat([o for o in out], 1) if isinstance(out, tuple) else out
        w = torch.randn(logits.shape, generator=torch.Generator().manual_seed(6)).cuda() * 1e-2
        (logits * w).sum().backward()
        torch.cuda.synchronize()
        assert eng.workspace(7, True).pruned == (mode == "rows")
        outs[mode] = (logits.detach().float().cpu(), {n_: p.grad.detach().float().cpu().clone() for n_, p in m.named_parameters()
                                                       if p.grad is not None})
    la, ga = outs["rows"]
    lb, gb = outs["dense"]
    assert torch.equal(la, lb)                                      # the same rows through the same kernels: bit-identical logits
    assert set(ga) == set(gb)
    for n_ in gb:
        den = gb[n_].norm().item()
        if den == 0.0:
            assert ga[n_].abs().max().item() == 0.0, n_
        else:
            assert (ga[n_] - gb[n_]).norm().item() / den < 2e-3, (n_, (ga[n_] - gb[n_]).norm().item() / den)


@pytest.mark.gpu
@pytest.mark.parametrize("N,H,q_rows", [(198, 3, 2), (197, 12, 1), (64, 2, 2), (198, 3, 70), (198, 3, 130), (577, 3, 2), (577, 2, 130),
                                         (1025, 1, 1)])
def test_attention_leading_query_rows(N, H, q_rows):
    """vitk_attention_fwd / _bwd with q_rows: the leading rows of out / lse equal the full forward's, and with dout zero from
    row q_rows on the backward's dqkv equals the full backward's (bit-exact: the skipped chunks only ever add zeros)."""
    B = 5
    g = torch.Generator().manual_seed(N + H)
    qkv = (torch.randn(B, N, 3 * H * 64, generator=g) * 0.5).cuda().half()
    scale = 64 ** -0.5
    out_f, lse_f = ops.attention_fwd(qkv, B, N, H, scale)
    out_p = torch.full_like(out_f, float("nan"))
    lse_p = torch.full_like(lse_f, float("nan"))
    ops.attention_fwd(qkv, B, N, H, scale, out=out_p, lse=lse_p, q_rows=q_rows)
    assert torch.equal(out_p[:, :q_rows], out_f[:, :q_rows]) and torch.equal(lse_p[:, :, :q_rows], lse_f[:, :, :q_rows])
    dout = torch.zeros_like(out_f)
    dout[:, :q_rows] = (torch.randn(B, q_rows, H * 64, generator=g) * 0.1).cuda().half()
    full = ops.attention_bwd(qkv, out_f, dout, lse_f, B, N, H, scale)
    part = ops.attention_bwd(qkv, out_p, dout, lse_p, B, N, H, scale, q_rows=q_rows)   # unwritten rows of out / lse hold NaNs
    torch.cuda.synchronize()
    assert torch.isfinite(part.float()).all()
    assert torch.equal(part, full)


# ------------------------------------------------------------------ class-token heat map (attention_utils.py:50-67)
HEAT_TOL = 2e-6   # fp32 bilinear blend of values <= 1: summation order / fma contraction only


def test_cls_attention_heatmap_matches_reference_fixture_and_oracle():
    """vitk_cls_attention_heatmap against (a) the maps the reference's own visualize_attention_maps drew
    (tests/golden/cls_heatmap.pt) and (b) the oracle on other shapes: 5-D / 4-D maps, a strided view of the image-major
    buffer the ensemble keeps, a rollout matrix, a rollout row with two prefix tokens, a ready grid, odd output widths."""
    rec = torch.load(ROOT / "tests" / "golden" / "cls_heatmap.pt")
    small, full = rec["small_maps"].cuda(), rec["full_maps"].cuda()
    assert (ENS.cls_attention_heatmap(small, 32).cpu() - rec["small_32x32_last"]).abs().max().item() < HEAT_TOL
    assert (ENS.cls_attention_heatmap(small, (24, 40), layer_idx=0).cpu() - rec["small_24x40_first"]).abs().max().item() < HEAT_TOL
    assert (ENS.cls_attention_heatmap(full, (224, 224)).cpu() - rec["full_224"]).abs().max().item() < HEAT_TOL
    g = torch.Generator().manual_seed(9)
    maps = torch.randn(3, 5, 6, 38, 38, generator=g).softmax(-1)                # 36 patches + 2 prefix tokens (DeiT layout)
    for hw, layer in (((96, 96), -1), ((45, 67), 1), ((7, 9), 0)):               # up, odd sizes (scalar stores), down
        want = O.cls_attention_heatmap(maps, hw, layer, n_prefix=2)
        got = ENS.cls_attention_heatmap(maps.cuda(), hw, layer_idx=layer, n_prefix=2)
        assert got.shape == want.shape and (got.cpu() - want).abs().max().item() < HEAT_TOL
        got4 = ENS.cls_attention_heatmap(maps[layer].cuda(), hw, n_prefix=2)     # one layer's [B,H,N,N]
        assert torch.equal(got4, got)
        im = maps.transpose(0, 1).contiguous().cuda()                            # [B,L,H,N,N]: the ensemble's buffer
        assert torch.equal(ops.cls_attention_heatmap(im[:, layer], hw, n_prefix=2), got)   # strided batch, no copy
    # rollout matrix [B,N,N] -> its class-token row; a row [B,N]; a grid [B,g,g]
    roll = ops.attention_rollout(maps.cuda().contiguous(), "mean")
    want = torch.nn.functional.interpolate(roll[:, 0, 2:].reshape(5, 1, 6, 6).cpu(), size=(48, 48), mode="bilinear",
                                           align_corners=False)[:, 0]
    for src in (roll, roll[:, 0, :].contiguous()):
        assert (ENS.cls_attention_heatmap(src, 48, n_prefix=2).cpu() - want).abs().max().item() < HEAT_TOL
    assert (ENS.cls_attention_heatmap(roll[:, 0, 2:].reshape(5, 6, 6), 48, grid=True).cpu() - want).abs().max().item() < HEAT_TOL
    with pytest.raises(ValueError):
        ENS.cls_attention_heatmap(roll[:, :5, :7], 48)                           # 3-D but not square: neither a rollout matrix nor a grid
    # the reference's hard-wired `[0, 1:]` on a 38-token sequence leaves 37 columns: not a square grid -> ValueError here
    with pytest.raises(ValueError):
        ENS.cls_attention_heatmap(maps.cuda(), 48, n_prefix=1)
    with pytest.raises(RuntimeError):
        ENS.cls_attention_heatmap(maps, 48, n_prefix=2)                          # CPU tensor: no fallback


def test_cls_attention_heatmap_from_the_model_api_batch_256():
    """DeiT-tiny eval forward -> blocks[-1].attn.attention_maps -> heat maps for 256 images in one launch; properties that hold at
    any size: inside the grid's range, mean preserved by the 16x integer upscale, identical images give identical maps."""
    cfg = O.DEIT_TINY
    m, _ = build(cfg, 42)
    m.eval()
    m.store_attention = True
    x, _ = O.seeded_batch(cfg, 4, 42)
    x = x.repeat(64, 1, 1, 1)                                                    # 256 images, 64 copies of 4
    with torch.no_grad():
        m(x.cuda())
    last = m.blocks[-1].attn.attention_maps                                      # [B,H,N,N]; on the host like the reference's (:187-188)
    assert tuple(last.shape) == (256, 3, 198, 198)
    last = last.cuda()
    heat = ENS.cls_attention_heatmap(last, 224, n_prefix=2)
    torch.cuda.synchronize()
    assert heat.shape == (256, 224, 224)
    grid = last.float().mean(1)[:, 0, 2:].reshape(256, 14, 14)
    assert (heat.amin((1, 2)) >= grid.amin((1, 2)) - 1e-7).all() and (heat.amax((1, 2)) <= grid.amax((1, 2)) + 1e-7).all()
    assert ((heat.mean((1, 2)) - grid.mean((1, 2))).abs() < 1e-4 * grid.mean((1, 2))).all()
    assert (heat[:4] - heat[4:8]).abs().max().item() < 1e-6 and (heat[:4] - heat[252:]).abs().max().item() < 1e-6
    want = O.cls_attention_heatmap(last[:4].float().cpu().unsqueeze(0), (224, 224), -1, n_prefix=2)
    assert (heat[:4].cpu() - want).abs().max().item() < HEAT_TOL


# ------------------------------------------------------------------ long sequences through the whole model (tcgen05 key-tile kernels)
@pytest.mark.parametrize("img,patch,deit,batch", [(384, 16, True, 3), (384, 16, False, 2), (224, 8, False, 2)])
def test_long_sequence_models_vs_oracle(img, patch, deit, batch):
    """384x384 images (577 / 578 tokens; tests/test_vision_transformer_base.py:309-319 exercises this size) and patch 8 at 224
    (785 tokens): logits, loss and every parameter gradient of a 2-block model against the oracle.  These shapes run the
    key-tile forward and the streaming dK/dV + dQ kernels, with the last block pruned to the classifier's rows (q_rows)."""
    cfg = O.VitConfig(img_size=img, patch_size=patch, embed_dim=128, depth=2, num_heads=2, is_deit=deit, distilled=deit)
    model, sd = build(cfg, 31)
    x, y = O.seeded_batch(cfg, batch, 31)
    loss, outs, grads = run_gpu(model, x, y)
    ref_loss, ref_out, ref_grads = O.train_step(sd, x, y, cfg)
    ref_outs = ref_out if isinstance(ref_out, tuple) else (ref_out,)
    for o, ref in zip(outs, ref_outs):
        assert (o - ref.detach()).abs().max().item() < LOGIT_TOL
    assert abs(loss - ref_loss.item()) < 2e-3
    worst = max((rel_l2(grads[n], g), n) for n, g in ref_grads.items() if g is not None)
    assert worst[0] < GRAD_TOL, worst
    model.eval()
    with torch.no_grad():
        ev = model(x.cuda())
    ref_ev = O.forward(sd, x, cfg, training=False)
    assert (ev.cpu() - ref_ev).abs().max().item() < LOGIT_TOL
    maps = model.blocks[-1].attn.attention_maps                       # fp32 maps beyond one S tile: the CUDA-core row kernel
    assert maps.shape[-1] == cfg.num_patches + cfg.num_prefix and (maps.sum(-1) - 1).abs().max().item() < 1e-5
