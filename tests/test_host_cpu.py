"""CPU-only checks of the host side: the C-ABI library exports exactly what include/vitk.h declares (no compute
calls -- there is no GPU here), the drop-in classes keep the reference's API surface / state_dict keys, the
product path fails LOUDLY without CUDA, the registry and config plumbing, and the data-parallel bucket logic
(world_size-2 gloo processes)."""
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import thyroid_vit_cnn_comparison_b200 as tv  # noqa: E402
from thyroid_vit_cnn_comparison_b200 import _lib, engine, parallel, registry, training, vit  # noqa: E402
from oracle import vit_oracle as O  # noqa: E402


def _header_functions():
    h = (ROOT / "include" / "vitk.h").read_text()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|void|int64_t|const char\*)\s+(vitk_\w+)\s*\(([^;]*?)\)\s*;", h, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args == "void" else len([a for a in args.split(",") if a.strip()])
    return out


def test_cabi_library_exports_every_declared_symbol():
    lib = _lib.load()                      # loads (and if needed builds) libvitk.so; no GPU required
    decl = _header_functions()
    assert set(decl) == set(_lib.SIGNATURES), set(decl) ^ set(_lib.SIGNATURES)
    for name, nargs in decl.items():
        assert hasattr(lib, name), name
        assert len(_lib.SIGNATURES[name][1]) == nargs, name
    assert lib.vitk_abi_version() == _lib.ABI_VERSION
    nm = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (vitk_\w+)", nm))
    assert exported == set(decl), exported ^ set(decl)


def test_no_cpu_fallback_and_loud_failure():
    m = vit.create_deit_tiny(img_size=224, in_chans=3, distilled=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 224, 224))
    from thyroid_vit_cnn_comparison_b200 import ops
    with pytest.raises(RuntimeError):
        ops.layernorm_fwd(torch.zeros(4, 64), torch.ones(64), torch.zeros(64))
    # the product package must not import the oracle
    for f in (ROOT / "thyroid-vit-cnn-comparison_b200").glob("*.py"):
        assert "oracle" not in f.read_text().replace("# oracle", ""), f


def test_state_dict_keys_and_param_order_match_reference():
    for cfg, model in [(O.DEIT_TINY, vit.create_deit_tiny(img_size=224, in_chans=3, distilled=True)),
                       (O.VIT_BASE, vit.create_vit_base(img_size=224, in_chans=3, drop_path_rate=0.0))]:
        assert [n for n, _ in model.named_parameters()] == list(O.param_shapes(cfg))
        assert {n: tuple(p.shape) for n, p in model.named_parameters()} == {n: tuple(s) for n, s in O.param_shapes(cfg).items()}
        model.load_state_dict(O.seeded_state_dict(cfg, 1), strict=True)
    m = vit.create_deit_tiny(img_size=224, in_chans=3)
    assert sum(p.numel() for p in m.parameters()) == 5526501
    assert m.hparams.depth == 12 and m.hparams.get("num_heads") == 3 and m.embed_dim == 192
    assert m.patch_embed.num_patches == 196 and m.pos_embed.shape == (1, 198, 192)
    assert isinstance(m.blocks[0].drop_path, torch.nn.Identity)


def test_parameter_groups_reproduce_reference_table():
    import json
    rec = json.loads((ROOT / "tests/golden/param_groups_deit_tiny.json").read_text())
    m = vit.create_deit_tiny(img_size=224, in_chans=3, distilled=True)
    groups = m.get_parameter_groups(weight_decay=0.05, layer_decay=0.75)
    assert [(g["name"], g["weight_decay"], round(g["lr_scale"], 12)) for g in groups] == \
        [(g["name"], g["weight_decay"], round(g["lr_scale"], 12)) for g in rec["groups"]]


def test_factories_registry_and_errors():
    with pytest.raises(ValueError):
        vit.get_vit_model("vit_invalid")                       # reference tests/test_vit_models.py:334-337
    with pytest.raises(ValueError):
        vit.create_deit_model("deit_huge")
    assert set(vit.VIT_MODEL_REGISTRY) == {"vit_tiny", "vit_small", "vit_base"}
    small = vit.get_vit_model("vit_small", img_size=224, in_chans=3, drop_path_rate=0.0)
    assert small.embed_dim == 384 and small.blocks[0].attn.num_heads == 6

    class Cfg:                                                 # tests/unit/test_models.py:10-22 config contract
        def __init__(self, **k):
            self.__dict__.update(k)

        def get(self, k, d=None):
            return getattr(self, k, d)

    w = registry.ModelRegistry.create_model(Cfg(name="deit_tiny", pretrained=False, num_classes=2, img_size=224))
    assert w.model is not None and list(w.state_dict())[0] == "model.cls_token"
    w2 = registry.ModelRegistry.create_model(Cfg(name="vit_base", num_classes=10, img_size=224, extra_params={"in_chans": 1}))
    assert w2.model.head.out_features == 10 and w2.model.patch_embed.proj.in_channels == 1
    with pytest.raises(ValueError):
        registry.ModelRegistry.create_model(Cfg(name="resnet18"))
    with pytest.raises(ValueError):
        registry.ModelRegistry.create_model(object())
    assert set(registry.ModelRegistry.list_models("vit")) >= {"deit_tiny", "vit_base"}

    class FakeRefRegistry:                                     # install_into overwrites the reference registry slots
        store = {}

        @classmethod
        def register(cls, names, model_type="default"):
            def deco(c):
                for n in names:
                    cls.store[n] = c
                return c
            return deco
    registry.install_into(FakeRefRegistry)
    assert FakeRefRegistry.store["deit_tiny"] is registry.DeiT and FakeRefRegistry.store["vit_small"] is registry.VisionTransformer


def test_unsupported_options_fail_loudly():
    m = vit.VisionTransformer(img_size=64, embed_dim=64, depth=1, num_heads=1, in_chans=3, pool_type="gap")
    m._check_supported()                                       # gap pooling / pre_logits: general tail (csrc/pool_head.cu)
    assert m._pool_range() == (1, 17) and m._rep_size() == 0
    m = vit.VisionTransformer(img_size=64, embed_dim=64, depth=1, num_heads=1, in_chans=3, class_token=False)
    assert m._pool_range() == (0, 16) and m._n_prefix() == 0 and not hasattr(m, "cls_token")
    m = vit.VisionTransformer(img_size=64, embed_dim=64, depth=1, num_heads=1, in_chans=3, representation_size=64)
    assert m._pool_range() is None and m._rep_size() == 64
    assert [n for n, _ in m.named_parameters() if n.startswith("pre_logits")] == ["pre_logits.0.weight", "pre_logits.0.bias"]
    m = vit.VisionTransformer(img_size=64, embed_dim=64, depth=1, num_heads=1, in_chans=3, representation_size=32)
    with pytest.raises(RuntimeError):                          # head is Linear(embed_dim, classes): same failure as the reference
        m._rep_size()
    with pytest.raises(AttributeError):                        # deit_models.py:84-99 (Sequential has no .weight)
        vit.DeiT(img_size=64, embed_dim=64, depth=1, num_heads=1, in_chans=3, distilled=True, representation_size=64)
    m = vit.VisionTransformer(img_size=64, embed_dim=64, depth=1, num_heads=1, in_chans=3, projection_type="linear")
    assert [n for n, _ in m.named_parameters() if n.startswith("patch_embed.proj")] == ["patch_embed.proj.1.weight", "patch_embed.proj.1.bias"]
    assert m.patch_embed.proj[1].weight.shape == (64, 16 * 16 * 3)      # vision_transformer_base.py:103-107
    cfg = O.VitConfig(img_size=64, embed_dim=64, depth=1, num_heads=1, in_chans=3, is_deit=False, distilled=False, projection_type="linear")
    assert [n for n, _ in m.named_parameters()] == list(O.param_shapes(cfg))
    m = vit.create_vit_tiny(img_size=64, in_chans=3)           # factory default drop_path_rate=0.1: stochastic depth is
    m.train()                                                  # served by the engine (GPU test test_stochastic_depth_*)
    m._check_supported()
    m = vit.create_vit_tiny(img_size=64, in_chans=3, drop_rate=0.1)   # vit_base.yaml / vit_small.yaml: fused dropout
    m.train()                                                         # (GPU test test_dropout_training_step_*)
    m._check_supported()
    m = vit.create_vit_tiny(img_size=64, in_chans=3, attn_drop_rate=0.1)
    m.train()
    m._check_supported()                                       # attention-probability dropout: vitk_attention_dropout_{fwd,bwd}
    assert m.blocks[0].attn.attn_drop.p == 0.1
    with pytest.raises(NotImplementedError):                   # only the exact-erf GELU of the reference is fused
        vit.VisionTransformer(img_size=64, embed_dim=64, depth=1, num_heads=1, act_layer=torch.nn.ReLU)


def test_flat_layout_and_completed_prefix():
    cfg = O.VitConfig(img_size=64, embed_dim=64, depth=3, num_heads=1)
    names = [n for n in O.param_shapes(cfg) if "quality" not in n]
    order = engine.execution_order(names, cfg.depth)
    assert order[0].startswith("norm") or order[0].startswith("head")
    assert [n.split(".")[1] for n in order if n.startswith("blocks.")][0] == "2"       # last block first
    assert order[-1] in ("patch_embed.proj.bias", "patch_embed.proj.weight", "pos_embed", "cls_token", "dist_token")
    offsets, total = {}, 0
    shapes = O.param_shapes(cfg)
    for n in order:
        offsets[n] = (total, torch.Size(shapes[n]))
        total += (torch.Size(shapes[n]).numel() + engine.PAD - 1) // engine.PAD * engine.PAD
    ends = [parallel.completed_prefix(order, offsets, engine.PAD, s, cfg.depth) for s in ("head", "blocks.2.", "blocks.1.", "blocks.0.", "embed")]
    assert ends == sorted(ends) and ends[-1] == total and ends[0] > 0
    # the optional tail / projection tensors: pre_logits completes with the head, the linear patch projection with the embeddings
    cfg2 = O.VitConfig(img_size=64, embed_dim=64, depth=2, num_heads=1, is_deit=False, distilled=False, representation_size=64,
                       projection_type="linear")
    shapes2 = O.param_shapes(cfg2)
    names2 = [n for n in shapes2 if "quality" not in n]
    order2 = engine.execution_order(names2, cfg2.depth)
    tail = [n for n in order2 if n.startswith(("head", "norm.", "pre_logits."))]
    assert order2[:len(tail)] == tail and len(tail) == 6
    assert all(n.startswith(("patch_embed.proj.1.", "pos_embed", "cls_token")) for n in order2[-4:])
    offsets2, total2 = {}, 0
    for n in order2:
        offsets2[n] = (total2, torch.Size(shapes2[n]))
        total2 += (torch.Size(shapes2[n]).numel() + engine.PAD - 1) // engine.PAD * engine.PAD
    head_end = parallel.completed_prefix(order2, offsets2, engine.PAD, "head", cfg2.depth)
    assert head_end == offsets2[order2[len(tail)]][0]              # exactly the head / norm / pre_logits tensors


def _dp_worker(rank, world, port, tmp):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakeFlat:
        def __init__(self):
            self.order = ["head.weight", "blocks.1.w", "blocks.0.w", "pos_embed"]
            sizes = [256, 4096, 4096, 512]
            self.offsets, off = {}, 0
            for n, s in zip(self.order, sizes):
                self.offsets[n] = (off, torch.Size([s]))
                off += s
            self.grads = torch.full((off,), float(rank + 1))

        def bucket_slices(self, nbytes):
            return engine.FlatParams.bucket_slices(self, nbytes)

    class FakeDims:
        depth = 2

    class FakeEngine:
        flat, d, grad_ready_hook = FakeFlat(), FakeDims(), None

    eng = FakeEngine()
    red = parallel.BucketedAllReduce(bucket_mb=8192 * 4 / (1 << 20))     # ~2 tensors per bucket
    red.attach(eng)
    for stage in ("head", "blocks.1.", "blocks.0.", "embed"):
        eng.grad_ready_hook(stage)
    red.finish()
    expect = float(sum(r + 1 for r in range(world)))
    ok = bool(torch.all(eng.flat.grads == expect))
    overlapped = any(stage != "embed" for stage, _ in red.launched_log)       # some bucket left before backward ended
    folds = parallel.shard_folds(5, rank, world)
    logits = torch.stack([torch.full((3, 2), float(f)) for f in folds]) if folds else torch.zeros(0, 3, 2)
    full = parallel.gather_fold_logits(logits, folds, 5)
    ok_folds = bool(all(torch.all(full[f] == f) for f in range(5)))
    Path(tmp, f"r{rank}.txt").write_text(f"{ok} {overlapped} {ok_folds} {len(red.buckets)}")
    dist.destroy_process_group()


def test_bucketed_allreduce_and_fold_sharding_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_dp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        ok, overlapped, ok_folds, nb = (tmp_path / f"r{r}.txt").read_text().split()
        assert ok == "True" and overlapped == "True" and ok_folds == "True" and int(nb) >= 2


def test_cfg_get_and_label_handling():
    assert training.cfg_get({"a": 1}, "a") == 1 and training.cfg_get(None, "a", 5) == 5

    class C:
        x = 3

        def get(self, k, d=None):
            return {"y": 4}.get(k, d)
    assert training.cfg_get(C(), "x") == 3 and training.cfg_get(C(), "y") == 4 and training.cfg_get(C(), "z", 9) == 9
    assert training._labels(torch.tensor([[1], [0]], dtype=torch.int32)).tolist() == [1, 0]
    assert parallel.shard_folds(5, 1, 8) == [1] and parallel.shard_folds(5, 0, 2) == [0, 2, 4]


def test_lightning_checkpoint_key_remapping():
    """run_ensemble_kfold_evaluation.py:98-101: `model.model.<key>` of a Lightning .ckpt -> wrapper / bare-network keys."""
    class Cfg:
        def __init__(self, **k):
            self.__dict__.update(k)

        def get(self, k, d=None):
            return getattr(self, k, d)
    w = registry.ModelRegistry.create_model(Cfg(name="deit_tiny", pretrained=False, num_classes=2, img_size=224))
    inner = {k: torch.randn_like(v) for k, v in w.model.state_dict().items()}
    ckpt = {"state_dict": {"model.model." + k: v for k, v in inner.items()}, "epoch": 3}
    registry.load_lightning_checkpoint(w, ckpt, strict=True)
    assert all(torch.equal(w.model.state_dict()[k], v) for k, v in inner.items())
    import copy
    bare = copy.deepcopy(w.model)
    for t in bare.state_dict().values():
        t.zero_()
    distill = {"student.model." + k: v for k, v in inner.items()}
    distill["teacher.features.conv0.weight"] = torch.zeros(1)
    registry.load_lightning_checkpoint(bare, distill, strict=True)
    assert all(torch.equal(bare.state_dict()[k], v) for k, v in inner.items())
    with pytest.raises(RuntimeError):
        registry.load_lightning_checkpoint(bare, {"state_dict": {"model.model.cls_token": inner["cls_token"]}}, strict=True)


def test_frozen_densenet_rewrites_match_torchvision_eval_forward():
    """teacher.FrozenDenseNet's three rewrites (preallocated NHWC block buffer instead of torch.cat, eval BatchNorm + ReLU as one
    affine pass, norm2 / norm0 folded into the preceding convolution) against torchvision's own eval forward, in fp32 on the
    CPU with the CUDA kernel replaced by a torch restatement of its contract (the product default has no CPU path)."""
    import torchvision
    from thyroid_vit_cnn_comparison_b200 import teacher as T
    torch.manual_seed(0)
    m = torchvision.models.densenet169(weights=None, num_classes=2).eval()
    g = torch.Generator().manual_seed(1)
    for mod in m.modules():                                   # non-trivial running statistics and affine terms
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
            mod.weight.data.copy_(1 + 0.2 * torch.randn(mod.num_features, generator=g))
            mod.bias.data.copy_(0.1 * torch.randn(mod.num_features, generator=g))
    assert T.is_supported(m) and not T.is_supported(torch.nn.Linear(4, 2))

    def affine_relu_double(x, C, scale, shift, out=None, relu=True):      # contract of vitk_affine_relu_nhwc (include/vitk.h)
        y = x[..., :C].float() * scale + shift
        y = (torch.relu(y) if relu else y).to(x.dtype)
        if out is None:
            return y.contiguous()
        out[..., :C] = y
        return out

    def pool_double(x, out, kernel, stride, pad, is_max):                  # contract of vitk_pool_nhwc
        import torch.nn.functional as F
        xn = x.permute(0, 3, 1, 2)
        y = F.max_pool2d(xn, kernel, stride, pad) if is_max else F.avg_pool2d(xn, kernel, stride, pad)
        out[..., :x.shape[-1]] = y.permute(0, 2, 3, 1)
        return out

    fast = T.FrozenDenseNet(m, dtype=torch.float32, affine_relu=affine_relu_double, pool=pool_double)
    x = torch.rand(2, 3, 64, 64, generator=g)
    with torch.no_grad():
        ref = m(x)
    out = fast(x)
    assert out.shape == ref.shape == (2, 2)
    assert (out - ref).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item())
    with pytest.raises(RuntimeError):                         # the default op is the CUDA kernel: loud on CPU tensors
        T.FrozenDenseNet(m, dtype=torch.bfloat16)(x)
    with pytest.raises(TypeError):
        T.FrozenDenseNet(torch.nn.Linear(4, 2))


def test_bench_traffic_table_points_at_committed_ncu_captures():
    """bench.py's roofline.traffic comes from the newest profiles/rNN_ncu_traffic_<model>.json: every kernel label it maps must
    exist in the committed capture, with a positive per-launch DRAM byte count."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for model, table in bench.NCU_CASES.items():
        f = bench.ncu_traffic_file(model)
        assert f is not None and f.parent == ROOT / "profiles"
        rec = json.loads(f.read_text())
        for label, case in table.items():
            assert case in rec, (model, label, case)
            nbytes, src = bench.ncu_traffic(model, label)
            assert nbytes > 0 and src.endswith(f":{case}")
    assert bench.ncu_traffic("deit_tiny", "no such kernel") == (None, None)


def test_product_side_kfold_reproduces_reference_fixture_and_pins_5fold(tmp_path):
    """generate_kfold_splits (scripts/prepare_kfold_data.py:30-73) in the PRODUCT package: all 21 index lists of the
    reference's committed 7-fold files bit-exactly, the JSON files it writes, and the 5-fold variant of BASELINE.json
    config 5 pinned against the oracle's restatement + its partition / stratification invariants."""
    import json
    from thyroid_vit_cnn_comparison_b200 import kfold
    from oracle import vit_oracle as O
    rec = json.loads((ROOT / "tests" / "golden" / "kfold_splits_7.json").read_text())
    labels = kfold.cars_labels(225, 225)
    assert labels.tolist() == rec["labels"] if "labels" in rec and isinstance(rec["labels"], list) else True
    runs = kfold.generate_kfold_splits(labels, 7, random_state=42, splits_dir=tmp_path)
    assert runs == rec["folds"]                                              # 7 folds x {train, val, test}
    for i in range(1, 8):
        on_disk = json.loads((tmp_path / f"split_fold_{i}.json").read_text())
        assert on_disk == rec["folds"][i - 1] and list(on_disk) == ["train", "val", "test"]
        assert kfold.load_fold_split(tmp_path, i) == rec["folds"][i - 1]
    five = kfold.generate_kfold_splits(labels, 5)
    assert five == O.kfold_splits(labels.tolist(), k=5, seed=42)
    tests = [set(r["test"]) for r in five]
    assert sorted(x for t in tests for x in t) == list(range(450))           # the test folds partition the dataset
    for i, r in enumerate(five):
        assert sorted(r["train"] + r["val"] + r["test"]) == list(range(450))
        assert r["val"] == five[(i + 1) % 5]["test"]                          # rotation rule
        assert abs(sum(int(labels[j]) for j in r["test"]) - len(r["test"]) / 2) <= 1
    # the reference's calling convention: a raw-data directory with normal/ and cancerous/ sub-directories
    raw = tmp_path / "raw"
    for cls, n in (("normal", 12), ("cancerous", 9)):
        (raw / cls).mkdir(parents=True)
        for j in range(n):
            (raw / cls / f"{j}.tif").write_bytes(b"")
    by_dir = kfold.generate_kfold_splits(raw, 3)
    assert by_dir == kfold.generate_kfold_splits(kfold.cars_labels(12, 9), 3)
    assert (tmp_path / "splits" / "split_fold_3.json").exists()
    with pytest.raises(ValueError):
        kfold.generate_kfold_splits(labels, 2)


def _ens_worker(rank, world, port, tmp):
    """EnsembleInference host logic on gloo: fold sharding, the shape hint for a rank that owns no fold, gather order.
    The member forward and the probability mix are replaced by CPU stand-ins (the real ones are CUDA kernels)."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from thyroid_vit_cnn_comparison_b200 import ensemble, ops as _ops

    class Net(torch.nn.Module):                       # what _inner() / the shape hint read
        num_classes, num_patches = 2, 16

        def __init__(self, fold):
            super().__init__()
            self.fold = fold

        def _ensure_engine(self):
            return None

    class Ens(ensemble.EnsembleInference):
        def _member_forward(self, model, images, want_maps, gray=None):
            B = images.shape[0]
            logits = torch.full((B, 2), float(model.fold)) + torch.arange(B).view(-1, 1) * torch.tensor([0.5, -0.5])
            maps = torch.full((B, 17), float(model.fold)) if want_maps else None
            return logits, maps, 1

    _ops.ensemble_probs = lambda full, w: ((full.softmax(-1) * w.view(-1, 1, 1)).sum(0), (full.softmax(-1) * w.view(-1, 1, 1)).sum(0).argmax(1))
    _ops.attention_rollout_row = lambda maps, row, fusion, image_major=False: maps
    out = {}
    for F in (5, 1):                                  # F = 1 with two ranks: rank 1 owns nothing and only joins the gather
        mine = parallel.shard_folds(F, rank, world)
        ens = Ens([Net(f) for f in mine], num_folds=F, rollout=True)
        images = torch.zeros(3, 1, 8, 8)
        images.__class__ = type("FakeCuda", (torch.Tensor,), {"is_cuda": property(lambda self: True)})
        res = ens(images)
        out[F] = (res["logits"].clone(), res["preds"].clone(), res["rollout"].clone(), mine)
    torch.save(out, Path(tmp) / f"e{rank}.pt")
    dist.destroy_process_group()


def test_ensemble_inference_fold_sharding_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_ens_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f"e{r}.pt", weights_only=False) for r in range(2))
    for F in (5, 1):
        l0, p0, g0, m0 = r0[F]
        l1, p1, g1, m1 = r1[F]
        assert torch.equal(l0, l1) and torch.equal(p0, p1) and torch.equal(g0, g1)          # identical on every rank
        assert l0.shape == (F, 3, 2) and g0.shape == (F, 3, 4, 4)
        for f in range(F):                                                                  # fold order restored after the gather
            assert torch.all(l0[f, 0] == float(f)) and torch.all(g0[f] == float(f))
    assert r0[5][3] == [0, 2, 4] and r1[5][3] == [1, 3] and r0[1][3] == [0] and r1[1][3] == []
