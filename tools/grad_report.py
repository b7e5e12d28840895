"""Per-parameter gradient error of the CUDA path vs the oracle (diagnostic, run on the GPU box)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from oracle import vit_oracle as O
import test_model_gpu as T

def report(cfg, batch, seed, tag):
    model, sd = T.build(cfg, seed)
    x, y = O.seeded_batch(cfg, batch, seed)
    loss, outs, grads = T.run_gpu(model, x, y)
    dev = "cuda"
    sdg = {k: v.to(dev) for k, v in sd.items()}
    torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
    ref_loss, ref_out, ref_grads = O.train_step(sdg, x.to(dev), y.to(dev), cfg)
    ref_outs = ref_out if isinstance(ref_out, tuple) else (ref_out,)
    print(f"== {tag}: loss {loss:.6f} ref {ref_loss.item():.6f}; logits max-abs", max((o - r.detach().cpu()).abs().max().item() for o, r in zip(outs, ref_outs)))
    rows = sorted(((T.rel_l2(grads[n], g), n) for n, g in ref_grads.items() if g is not None), reverse=True)
    for e, n in rows[:12]:
        print(f"   {e:.4e}  {n}")
    print("   median", rows[len(rows) // 2][0])

report(O.VitConfig(img_size=64, patch_size=16, in_chans=1, embed_dim=128, depth=1, num_heads=2, distilled=False, is_deit=False), 2, 43, "small_vit")
report(O.VitConfig(img_size=64, embed_dim=64, depth=2, num_heads=1), 3, 42, "small_deit")
report(O.DEIT_TINY, 32, 42, "deit_tiny_b32")
report(O.VIT_BASE, 2, 42, "vit_base_b2")
report(O.VIT_BASE, 16, 42, "vit_base_b16")
