"""Tiny end-to-end workload for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_step.py`:
one DeiT-tiny train step at batch 4 (every kernel of the hot path at its real shapes, ragged last tiles included) and one
small ViT step with dropout + stochastic depth, then an eval forward with attention maps and feature extraction, then two
steps of the non-default constructor options (gap pooling, pre_logits, linear patch projection, attention dropout, no class token)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from thyroid_vit_cnn_comparison_b200 import optim, training, vit  # noqa: E402


def main():
    torch.manual_seed(0)
    m = vit.create_deit_tiny(img_size=224, patch_size=16, in_chans=3, num_classes=2, distilled=True).cuda().train()
    opt = optim.FusedAdamW(m, lr=1e-4, weight_decay=1e-5, max_grad_norm=1.0)
    step = training.TrainStep(m, opt, 4, mode="ce", use_graph=False)
    x = torch.rand(4, 3, 224, 224, device="cuda")
    y = torch.randint(0, 2, (4,), device="cuda")
    for _ in range(2):
        st = step(x, y)
    torch.cuda.synchronize()
    print("deit_tiny loss", float(st[0]))
    v = vit.VisionTransformer(img_size=64, patch_size=16, in_chans=3, num_classes=2, embed_dim=128, depth=2, num_heads=2,
                              drop_rate=0.1, drop_path_rate=0.2).cuda().train()
    opt2 = optim.FusedAdamW(v, lr=1e-4, weight_decay=1e-5, max_grad_norm=1.0)
    step2 = training.TrainStep(v, opt2, 8, mode="ce", use_graph=False)
    x2 = torch.rand(8, 3, 64, 64, device="cuda")
    y2 = torch.randint(0, 2, (8,), device="cuda")
    for _ in range(2):
        st = step2(x2, y2)
    torch.cuda.synchronize()
    print("small vit (dropout) loss", float(st[0]))
    v.eval()
    with torch.no_grad():
        out = v(x2)
        feats = v.extract_features(x2)
    torch.cuda.synchronize()
    print("eval", tuple(out.shape), tuple(feats.shape), v.get_attention_maps().shape)
    # the non-default constructor options: gap pooling + pre_logits (pool_head.cu), linear patch projection (patchify_hwc),
    # attention-probability dropout (DROP instantiations of the tcgen05 attention kernels), no class token
    for kw in (dict(pool_type="gap", representation_size=128, projection_type="linear", attn_drop_rate=0.1, drop_rate=0.1),
               dict(class_token=False, attn_drop_rate=0.2)):
        g = vit.VisionTransformer(img_size=64, patch_size=16, in_chans=3, num_classes=2, embed_dim=128, depth=2, num_heads=2,
                                  drop_path_rate=0.1, **kw).cuda().train()
        opt3 = optim.FusedAdamW(g, lr=1e-4, weight_decay=1e-5, max_grad_norm=1.0)
        step3 = training.TrainStep(g, opt3, 8, mode="ce", use_graph=False)
        for _ in range(2):
            st = step3(x2, y2)
        torch.cuda.synchronize()
        print("options", sorted(kw), "loss", float(st[0]))


if __name__ == "__main__":
    main()
