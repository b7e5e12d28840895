"""Times vitk_attention_fwd / _bwd on long sequences (384x384 images: 577 tokens; patch 8: 785 / 1025), L2 flushed between
launches, CUDA events on the launching stream.  Usage: python tools/attn_long_bench.py [B N H]..."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import thyroid_vit_cnn_comparison_b200  # noqa: F401,E402
from thyroid_vit_cnn_comparison_b200 import ops  # noqa: E402


def bench(fn, flush, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    if "--once" in sys.argv:          # one forward + one backward launch (for an ncu capture): B N H follow the flag
        sys.argv.remove("--once")
        B, N, H = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else (32, 577, 12)
        g = torch.Generator().manual_seed(1)
        qkv = (torch.randn(B, N, 3 * H * 64, generator=g) * 0.5).cuda().half()
        dout = (torch.randn(B, N, H * 64, generator=g) * 0.1).cuda().half()
        out, lse = ops.attention_fwd(qkv, B, N, H, 0.125)
        ops.attention_bwd(qkv, out, dout, lse, B, N, H, 0.125)
        torch.cuda.synchronize()
        return
    shapes = [(32, 577, 12), (64, 577, 3), (16, 785, 12), (8, 1025, 12)]
    if len(sys.argv) > 3:
        a = [int(v) for v in sys.argv[1:]]
        shapes = [tuple(a[i:i + 3]) for i in range(0, len(a) - 2, 3)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    for B, N, H in shapes:
        g = torch.Generator().manual_seed(1)
        qkv = (torch.randn(B, N, 3 * H * 64, generator=g) * 0.5).cuda().half()
        dout = (torch.randn(B, N, H * 64, generator=g) * 0.1).cuda().half()
        out, lse = ops.attention_fwd(qkv, B, N, H, 0.125)
        dqkv = torch.empty_like(qkv)
        delta = torch.empty_like(lse)
        t_f = bench(lambda: ops.attention_fwd(qkv, B, N, H, 0.125, out=out, lse=lse), flush)
        t_b = bench(lambda: ops.attention_bwd(qkv, out, dout, lse, B, N, H, 0.125, dqkv=dqkv, delta=delta), flush)
        fl_f = 4.0 * B * H * N * N * 64
        rows.append({"B": B, "N": N, "H": H, "fwd_us": round(t_f, 1), "bwd_us": round(t_b, 1), "fwd_tflops": round(fl_f / t_f * 1e-6, 1),
                     "bwd_tflops": round(2.5 * fl_f / t_b * 1e-6, 1)})
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
