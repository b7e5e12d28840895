"""Where an ensemble member's eval step goes: encoder forward without maps, with fp32 attention-map emission, and the
class-token rollout (CUDA events, DeiT-tiny, batch 256)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import thyroid_vit_cnn_comparison_b200  # noqa
from thyroid_vit_cnn_comparison_b200 import ops, vit
from thyroid_vit_cnn_comparison_b200.engine import GraySpec

B = 256
name = sys.argv[1] if len(sys.argv) > 1 else "deit_tiny"
m = (vit.create_deit_tiny(img_size=224, in_chans=3, distilled=True) if name == "deit_tiny"
     else vit.create_vit_base(img_size=224, in_chans=3, drop_path_rate=0.0)).cuda().eval()
eng = m._ensure_engine()
d = eng.d
tiles = torch.randint(0, 65536, (B, 224, 224), dtype=torch.int32).to(torch.uint16).cuda()
maps = torch.empty(B, d.depth, d.heads, d.tokens, d.tokens, device="cuda")   # image-major


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    t_plain = timed(lambda: eng.forward(tiles, train=False, gray=GraySpec()))
    t_maps = timed(lambda: eng.forward(tiles, train=False, attn_probs=("image_major", maps), gray=GraySpec()))
    t_roll = timed(lambda: ops.attention_rollout_row(maps, 0, "mean", image_major=True))
    qkv = torch.randn(B, d.tokens, 3 * d.dim, device="cuda").to(torch.float16)
    pr = torch.empty(B, d.heads, d.tokens, d.tokens, device="cuda")
    t_attn_maps = timed(lambda: ops.attention_fwd(qkv, B, d.tokens, d.heads, 0.125, probs=pr))
    t_attn = timed(lambda: ops.attention_fwd(qkv, B, d.tokens, d.heads, 0.125))
print(f"{name} B={B}: eval forward {t_plain:.2f} ms | with map emission {t_maps:.2f} ms | rollout row {t_roll:.2f} ms | "
      f"one attention layer: {t_attn * 1e3:.0f} us, with maps {t_attn_maps * 1e3:.0f} us (map bytes {pr.numel() * 4 / 1e6:.0f} MB)")
