"""TEST INFRASTRUCTURE -- generates tests/golden/* by running the REFERENCE's own classes
(imported unmodified from /root/reference via oracle/ref_loader.py) on seeded inputs.

Run in the authoring container only:   python oracle/make_golden.py
The fixtures it writes are committed; the GPU box never sees /root/reference.

Fixtures
  small_deit.pt / small_vit.pt   tiny configs (DeiT: embed 64, 1 head, depth 2; ViT: embed 128, 2 heads, depth 1; 64x64 image): inputs are
                                 regenerated from seeds; stored = logits, loss, FULL gradients, parameters
                                 after one clip+AdamW step, eval-mode logits, attention maps of layer 0.
  deit_tiny_b4.pt / vit_base_b2.pt  full-size models: logits, loss, per-parameter gradient norm and the
                                 gradient's projection on a seeded random direction (compact but sensitive).
  small_vit_attn_dropout.pt      the same with attn_drop_rate = 0.2 on top (Attention.attn_drop, :184): its keep masks [B,H,N,N] per block too.
  small_vit_{gap_rep,nocls,cls_rep,linear_proj}.pt   constructor options outside the default tail / patch projection.
  small_vit_dropout.pt           ViT (embed 128, 2 heads, depth 2) with drop_rate = 0.1 in training mode: the keep masks every
                                 nn.Dropout call drew (forward hooks, site order pos_drop, then per block proj_drop, Mlp.drop #1,
                                 Mlp.drop #2), logits, loss, full gradients -- pins WHERE the reference applies dropout.
  ingest.pt                      input pipeline: raw uint16 tiles -> CARSThyroidDataset._preprocess_image (cv2.resize + /65535),
                                 AdaptiveNormalization('percentile'), MixUp / CutMix with recorded host draws -- outputs of the
                                 reference's own classes (src/data/{dataset,quality_preprocessing,vit_transforms}.py).
  param_groups_deit_tiny.json    get_parameter_groups() table (name, weight_decay, lr_scale) + named_parameters order.
  distill_loss.pt                DistillationLoss / training_step arithmetic on random logits.
  kfold_splits_7.json            the reference's committed data/splits/split_fold_{1..7}.json (known-answer vectors).
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_loader, vit_oracle as O  # noqa: E402

GOLD = ROOT / "tests" / "golden"

SMALL_DEIT = O.VitConfig(img_size=64, patch_size=16, in_chans=3, embed_dim=64, depth=2, num_heads=1, distilled=True, is_deit=True)
SMALL_VIT = O.VitConfig(img_size=64, patch_size=16, in_chans=1, embed_dim=128, depth=1, num_heads=2, distilled=False, is_deit=False)


def build_reference(cfg: O.VitConfig, seed: int):
    base, vitm, deit = ref_loader.load()
    kw = dict(img_size=cfg.img_size, patch_size=cfg.patch_size, in_chans=cfg.in_chans, num_classes=cfg.num_classes,
              embed_dim=cfg.embed_dim, depth=cfg.depth, num_heads=cfg.num_heads, mlp_ratio=cfg.mlp_ratio)
    if cfg.is_deit:
        model = deit.DeiT(distilled=cfg.distilled, **kw)
    else:
        opt = {}
        if not cfg.class_token:
            opt["class_token"] = False
        if cfg.pool_type != "cls":
            opt["pool_type"] = cfg.pool_type
        if cfg.representation_size:
            opt["representation_size"] = cfg.representation_size
        if cfg.projection_type != "conv":
            opt["projection_type"] = cfg.projection_type
        model = vitm.VisionTransformer(drop_path_rate=0.0, **kw, **opt)
    sd = O.seeded_state_dict(cfg, seed)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return model


def direction(name_idx: int, shape, seed: int = 1234):
    g = torch.Generator().manual_seed(seed * 7 + name_idx)
    return torch.randn(shape, generator=g)


def run_case(cfg: O.VitConfig, batch: int, seed: int, full: bool):
    torch.manual_seed(0)
    model = build_reference(cfg, seed)
    x, y = O.seeded_batch(cfg, batch, seed)
    model.train()
    out = model(x)
    if isinstance(out, tuple):
        loss = 0.5 * F.cross_entropy(out[0], y) + 0.5 * F.cross_entropy(out[1], y)   # lightning_modules.py:459-461
        logits = [o.detach().clone() for o in out]
    else:
        loss = F.cross_entropy(out, y)
        logits = [out.detach().clone()]
    loss.backward()
    names = [n for n, _ in model.named_parameters()]
    rec = {"config": cfg.__dict__, "batch": batch, "seed": seed, "loss": loss.item(), "logits": logits,
           "param_names": names}
    grads = {n: (p.grad.detach().clone() if p.grad is not None else None) for n, p in model.named_parameters()}
    rec["no_grad_params"] = [n for n, g in grads.items() if g is None]
    rec["grad_norm"] = {n: g.norm().item() for n, g in grads.items() if g is not None}
    rec["grad_proj"] = {n: (g * direction(i, g.shape)).sum().item() for i, (n, g) in enumerate(grads.items()) if g is not None}
    if full:
        rec["grads"] = {n: g for n, g in grads.items() if g is not None}
        model.eval()                                                                       # eval path BEFORE the update
        with torch.no_grad():
            rec["eval_logits"] = model(x).detach().clone()
        rec["attn_layer0"] = model.blocks[0].attn.attention_maps.clone()                   # vision_transformer_base.py:187-188
        if not cfg.is_deit:
            with torch.no_grad():
                rec["eval_features"] = model.extract_features(x).detach().clone()          # :494-497 (after pre_logits)
        model.train()
        # one optimizer step exactly as Lightning would run it: clip_grad_norm_(1.0) then AdamW
        params = [p for p in model.parameters() if p.grad is not None]
        groups_tbl = model.get_parameter_groups(weight_decay=0.05)
        for gdict in groups_tbl:
            gdict["lr"] = 1e-3 * gdict.get("lr_scale", 1.0)                                # lightning_modules.py:1101-1103
        groups_tbl = [g for g in groups_tbl if g["params"][0].grad is not None]
        opt = torch.optim.AdamW(groups_tbl, betas=(0.9, 0.999), weight_decay=0.05)           # :1108-1113
        total_norm = torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        rec["total_grad_norm"] = total_norm.item()
        rec["params_after_step"] = {n: p.detach().clone() for n, p in model.named_parameters()}
    return rec


# constructor options of the base class outside the default tail (vision_transformer_base.py:470-477, vit_models.py:97-106)
GAP_REP_VIT = O.VitConfig(img_size=64, patch_size=16, in_chans=3, embed_dim=128, depth=2, num_heads=2, distilled=False, is_deit=False,
                          pool_type="gap", representation_size=128)
NOCLS_VIT = O.VitConfig(img_size=64, patch_size=16, in_chans=1, embed_dim=64, depth=2, num_heads=1, distilled=False, is_deit=False,
                        class_token=False)
CLS_REP_VIT = O.VitConfig(img_size=64, patch_size=16, in_chans=1, embed_dim=64, depth=1, num_heads=1, distilled=False, is_deit=False,
                          representation_size=64)
LINEAR_VIT = O.VitConfig(img_size=64, patch_size=16, in_chans=3, embed_dim=64, depth=1, num_heads=1, distilled=False, is_deit=False,
                         projection_type="linear")
DROP_VIT = O.VitConfig(img_size=64, patch_size=16, in_chans=3, embed_dim=128, depth=2, num_heads=2, distilled=False, is_deit=False)
ATTN_DROP_VIT = O.VitConfig(img_size=64, patch_size=16, in_chans=3, embed_dim=128, depth=2, num_heads=2, distilled=False, is_deit=False)


def run_dropout_case(cfg: O.VitConfig, batch: int, seed: int, drop_rate: float, attn_drop_rate: float = 0.0):
    base, vitm, deit = ref_loader.load()
    kw = dict(img_size=cfg.img_size, patch_size=cfg.patch_size, in_chans=cfg.in_chans, num_classes=cfg.num_classes,
              embed_dim=cfg.embed_dim, depth=cfg.depth, num_heads=cfg.num_heads, mlp_ratio=cfg.mlp_ratio)
    model = vitm.VisionTransformer(drop_path_rate=0.0, drop_rate=drop_rate, attn_drop_rate=attn_drop_rate, **kw)
    sd = O.seeded_state_dict(cfg, seed)
    model.load_state_dict(sd, strict=True)
    calls = {}

    def hook(name):
        def fn(mod, inp, out):
            assert inp[0].ne(0).all()                 # so `out != 0` is exactly the keep mask
            calls.setdefault(name, []).append(out.detach().ne(0).to(torch.uint8))
        return fn
    model.pos_drop.register_forward_hook(hook("pos"))
    for i, blk in enumerate(model.blocks):
        blk.attn.proj_drop.register_forward_hook(hook(f"proj{i}"))
        blk.mlp.drop.register_forward_hook(hook(f"mlp{i}"))
        if attn_drop_rate > 0:
            blk.attn.attn_drop.register_forward_hook(hook(f"attn{i}"))                    # :184, input = softmax output (> 0)
    x, y = O.seeded_batch(cfg, batch, seed)
    torch.manual_seed(seed)
    model.train()
    out = model(x)
    loss = F.cross_entropy(out, y)
    loss.backward()
    masks = [calls["pos"][0]]
    for i in range(cfg.depth):
        assert len(calls[f"proj{i}"]) == 1 and len(calls[f"mlp{i}"]) == 2
        masks += [calls[f"proj{i}"][0], calls[f"mlp{i}"][0], calls[f"mlp{i}"][1]]
    attn_masks = [calls[f"attn{i}"][0] for i in range(cfg.depth)] if attn_drop_rate > 0 else None
    return {"config": cfg.__dict__, "batch": batch, "seed": seed, "drop_rate": drop_rate, "keep_masks": masks,
            "attn_drop_rate": attn_drop_rate, "attn_keep_masks": attn_masks,
            "loss": loss.item(), "logits": out.detach().clone(),
            "grads": {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}}


def _load_ref_file(rel: str, name: str):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, ref_loader.REF_ROOT / rel)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def run_ingest_case():
    """Runs the reference's own input-pipeline code on seeded raw tiles and records inputs + outputs."""
    import types
    import numpy as np
    stubs = str(ROOT / "oracle" / "_stubs")                     # tifffile (file I/O only) is absent here
    if stubs not in sys.path:
        sys.path.insert(0, stubs)
    qp = _load_ref_file("src/data/quality_preprocessing.py", "ref_quality_preprocessing")
    vt = _load_ref_file("src/data/vit_transforms.py", "ref_vit_transforms")
    # dataset.py imports src.config.schemas: resolve the reference's `src` package for this one import, then drop the
    # placeholder packages ref_loader may have installed under the same names
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "src" or k.startswith("src.")}
    sys.dont_write_bytecode = True
    sys.path.insert(0, str(ref_loader.REF_ROOT))
    try:
        import importlib
        ds = importlib.import_module("src.data.dataset")
    finally:
        sys.path.remove(str(ref_loader.REF_ROOT))
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    rng = np.random.default_rng(7)
    raws = []
    for (h, w) in [(300, 277), (224, 224), (256, 256), (100, 120)]:
        base = rng.integers(0, 65536, (h // 4 + 1, w // 4 + 1)).astype(np.float32)          # smooth-ish texture + noise
        img = np.kron(base, np.ones((4, 4), np.float32))[:h, :w] * 0.7 + rng.integers(0, 20000, (h, w))
        raws.append(np.clip(img, 0, 65535).astype(np.uint16))
    fake_self = types.SimpleNamespace(config=types.SimpleNamespace(img_size=224, channels=1))
    pre = [ds.CARSThyroidDataset._preprocess_image(fake_self, r.copy()) for r in raws]       # [1,224,224] float32 each
    batch = torch.stack(pre)                                                                 # [4,1,224,224]
    adaptive = qp.AdaptiveNormalization(method="percentile", percentiles=(1, 99))(batch.clone())
    u16 = torch.round(batch * 65535.0)
    assert torch.equal(u16 / 65535.0, batch)                                                 # exactly k / 65535: store k
    rec = {"raw": [torch.from_numpy(r.view(np.int16).copy()) for r in raws],                 # uint16 bit patterns
           "preprocessed_u16": torch.from_numpy(u16.to(torch.int32).numpy().astype(np.uint16).view(np.int16).copy()),
           "adaptive": adaptive}
    # MixUp / CutMix on a seeded normalised batch (regenerated from the seed by the tests); the host RNG is replayed so
    # that the draws can be recorded next to the outputs
    g = torch.Generator().manual_seed(99)
    images = torch.randn(4, 3, 64, 64, generator=g)
    labels = torch.tensor([0, 1, 1, 0])
    rec["mix_seed"], rec["labels"] = 99, labels
    np.random.seed(123); torch.manual_seed(123)
    lam = float(np.random.beta(0.8, 0.8)); index = torch.randperm(4)
    np.random.seed(123); torch.manual_seed(123)
    mixed, la, lb, lam_out = vt.MixUp(alpha=0.8)(images.clone(), labels)
    assert lam_out == lam and torch.equal(lb, labels[index])
    rec["mixup"] = {"lam": lam, "index": index, "out": mixed}
    np.random.seed(321); torch.manual_seed(321)
    lam_c = float(np.random.beta(1.0, 1.0)); index_c = torch.randperm(4)
    cx, cy = int(np.random.randint(64)), int(np.random.randint(64))
    np.random.seed(321); torch.manual_seed(321)
    cut, la, lb, lam_adj = vt.CutMix(alpha=1.0)(images.clone(), labels)
    assert torch.equal(lb, labels[index_c])
    rec["cutmix"] = {"lam_drawn": lam_c, "index": index_c, "cx": cx, "cy": cy, "out": cut, "lam": float(lam_adj)}
    return rec


def run_heatmap_case():
    """Class-token heat maps drawn by the reference's own visualize_attention_maps (attention_utils.py:14-81, run unmodified
    with a recording matplotlib stand-in, oracle/ref_loader.py) for seeded attention maps: grid 4 -> 32x32 and 24x40 images
    (layer -1 and layer 0), grid 14 -> 224x224 (DeiT-tiny's shape, ViT-style single class token)."""
    rec = {}
    g = torch.Generator().manual_seed(11)
    small = torch.randn(2, 3, 3, 17, 17, generator=g).softmax(-1)
    rec["small_maps"] = small
    rec["small_32x32_last"] = ref_loader.reference_cls_heatmaps(small, (32, 32), -1)
    rec["small_24x40_first"] = ref_loader.reference_cls_heatmaps(small, (24, 40), 0)
    full = torch.randn(1, 1, 3, 197, 197, generator=g).softmax(-1)
    rec["full_maps"] = full.to(torch.float32)
    rec["full_224"] = ref_loader.reference_cls_heatmaps(full, (224, 224), -1)
    return rec


def main():
    assert ref_loader.available(), "run this where /root/reference is mounted"
    GOLD.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    torch.save(run_heatmap_case(), GOLD / "cls_heatmap.pt")
    if "--heatmap-only" in sys.argv:
        return
    torch.save(run_case(GAP_REP_VIT, 3, 45, True), GOLD / "small_vit_gap_rep.pt")
    torch.save(run_case(NOCLS_VIT, 2, 46, True), GOLD / "small_vit_nocls.pt")
    torch.save(run_case(CLS_REP_VIT, 2, 47, True), GOLD / "small_vit_cls_rep.pt")
    torch.save(run_case(LINEAR_VIT, 2, 48, True), GOLD / "small_vit_linear_proj.pt")
    torch.save(run_dropout_case(ATTN_DROP_VIT, 2, 49, 0.1, 0.2), GOLD / "small_vit_attn_dropout.pt")
    if "--tail-only" in sys.argv:
        return
    torch.save(run_case(SMALL_DEIT, 3, 42, True), GOLD / "small_deit.pt")
    torch.save(run_case(SMALL_VIT, 2, 43, True), GOLD / "small_vit.pt")
    torch.save(run_dropout_case(DROP_VIT, 4, 44, 0.1), GOLD / "small_vit_dropout.pt")
    torch.save(run_ingest_case(), GOLD / "ingest.pt")
    torch.save(run_case(O.DEIT_TINY, 4, 42, False), GOLD / "deit_tiny_b4.pt")
    torch.save(run_case(O.VIT_BASE, 2, 42, False), GOLD / "vit_base_b2.pt")

    # parameter-group table of the real DeiT-tiny (incl. the substring quirk)
    model = build_reference(O.DEIT_TINY, 42)
    tbl = [{"name": g["name"], "weight_decay": g["weight_decay"], "lr_scale": g["lr_scale"]}
           for g in model.get_parameter_groups(weight_decay=0.05, layer_decay=0.75)]
    (GOLD / "param_groups_deit_tiny.json").write_text(json.dumps(
        {"named_parameters": [n for n, _ in model.named_parameters()], "groups": tbl,
         "num_params": sum(p.numel() for p in model.parameters())}, indent=0))

    # DistillationLoss on random logits (deit_models.py:461-480)
    _, _, deit = ref_loader.load()
    g = torch.Generator().manual_seed(7)
    c, d, t = (torch.randn(16, 2, generator=g) for _ in range(3))
    yy = torch.randint(0, 2, (16,), generator=g)
    recs = {"cls": c, "dist": d, "teacher": t * 2, "labels": yy, "cases": []}
    for kind in ("soft", "hard"):
        for ls in (0.0, 0.1):
            crit = deit.DistillationLoss(base_criterion=torch.nn.CrossEntropyLoss(label_smoothing=ls), distillation_type=kind,
                                         alpha=0.7, tau=3.0)
            recs["cases"].append({"type": kind, "label_smoothing": ls, "alpha": 0.7, "tau": 3.0,
                                  "loss": crit((c, d), yy, t * 2).item()})
    torch.save(recs, GOLD / "distill_loss.pt")

    # fold known-answer vectors
    folds = [json.loads((ref_loader.REF_ROOT / f"data/splits/split_fold_{i}.json").read_text()) for i in range(1, 8)]
    (GOLD / "kfold_splits_7.json").write_text(json.dumps(
        {"source": "reference data/splits/split_fold_{1..7}.json", "n": 450, "labels": "225 x 0 then 225 x 1",
         "folds": [{k: f[k] for k in ("train", "val", "test")} for f in folds]}))
    for f in sorted(GOLD.iterdir()):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
