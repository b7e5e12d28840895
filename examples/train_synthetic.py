"""End-to-end example on synthetic data: raw uint16 tiles -> GPU ingest -> captured training step -> on-device metrics.

    python examples/train_synthetic.py [--steps 300] [--batch 64]

Two classes of 96x96 uint16 "tiles": class 1 carries a faint bright blob at a random position on top of the same noise
texture as class 0.  A small ViT (the reference's class, libvitk underneath) has to reach > 95 % held-out accuracy within a
few hundred steps with dropout 0.1 + stochastic depth 0.1, fp16 operands, dynamic loss scaling, clip 1.0 and fused AdamW --
i.e. every piece of the hot path has to work together, not just agree with the oracle for one step."""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from thyroid_vit_cnn_comparison_b200 import ingest, metrics, optim, training, vit  # noqa: E402


def make_tiles(n: int, rng: np.random.Generator, size: int = 96):
    y = rng.integers(0, 2, n)
    img = rng.normal(20000.0, 3000.0, (n, size, size))
    yy, xx = np.mgrid[0:size, 0:size]
    for i in np.nonzero(y)[0]:
        cy, cx = rng.integers(16, size - 16, 2)
        img[i] += 9000.0 * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * 9.0 ** 2))
    return np.clip(img, 0, 65535).astype(np.uint16), y.astype(np.int64)


def u16_cuda(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(a.view(np.int16).copy()).view(torch.uint16).cuda()


def run(steps: int = 300, batch: int = 64, seed: int = 0, verbose: bool = True):
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    model = vit.VisionTransformer(img_size=64, patch_size=8, in_chans=3, num_classes=2, embed_dim=128, depth=3, num_heads=2,
                                  drop_rate=0.1, drop_path_rate=0.1).cuda().train()
    opt = optim.FusedAdamW(model, lr=1e-3, weight_decay=1e-4, max_grad_norm=1.0)
    step = training.TrainStep(model, opt, batch, mode="ce", use_graph=True)
    ing = ingest.TileIngest(64, 3, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), percentiles=(1, 99))
    losses = []
    for it in range(steps):
        raw, y = make_tiles(batch, rng)
        x = ing(u16_cuda(raw))                       # resize 96 -> 64, percentile normalisation, 3 channels, Normalize
        st = step(x, torch.from_numpy(y).cuda())
        if it % 20 == 0 or it == steps - 1:
            losses.append(float(st[0]))
            if verbose:
                print(f"step {it:4d}  loss {losses[-1]:.4f}  train_acc {float(st[3]) / batch:.3f}  loss_scale {float(model._engine.amp[0]):.0f}")
    model.eval()
    met = metrics.ClassificationMetrics(2)
    with torch.no_grad():
        for _ in range(8):
            raw, y = make_tiles(128, rng)
            met.update(model(ing(u16_cuda(raw))), torch.from_numpy(y).cuda())
    res = met.compute()
    if verbose:
        print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items() if k != "confusion"})
    return losses, res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--batch", type=int, default=64)
    a = ap.parse_args()
    run(a.steps, a.batch)
