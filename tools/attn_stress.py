"""Soak test of the tcgen05 attention kernels: tens of millions of CTAs of the forward / backward at the bench shapes,
every launch's result compared bit for bit with the first one (the kernels are deterministic).  It exists because the
backward once carried a barrier-phase race that deadlocked about one CTA in 1e7 (see the `bar_p` comment in
csrc/attention_tc.cu): unit tests cannot see such a rate, a soak of ~1e8 CTAs does.

    python tools/attn_stress.py [--model deit_tiny|vit_base] [--launches 100000] [--batch 256] [--check-every 2000]
"""
import argparse
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import thyroid_vit_cnn_comparison_b200  # noqa: E402,F401
from thyroid_vit_cnn_comparison_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="deit_tiny")
    ap.add_argument("--launches", type=int, default=100000)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--check-every", type=int, default=2000)
    ap.add_argument("--tokens", type=int, default=0, help="override the sequence length")
    a = ap.parse_args()
    T, H = (198, 3) if a.model == "deit_tiny" else (197, 12)
    if a.tokens:
        T = a.tokens
    B, D = a.batch, H * 64
    g = torch.Generator(device="cpu").manual_seed(0)
    qkv = torch.randn(B * T, 3 * D, generator=g).cuda().half()
    dout = (torch.randn(B * T, D, generator=g) * 0.1).cuda().half()
    scale = 64 ** -0.5
    out, lse = ops.attention_fwd(qkv, B, T, H, scale)
    ref_out = out.clone()
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(B, H, T, device="cuda")
    ops.attention_bwd(qkv, out, dout, lse, B, T, H, scale, dqkv=dqkv, delta=delta)
    ref_dqkv = dqkv.clone()
    torch.cuda.synchronize()
    t0 = time.time()
    done = 0
    while done < a.launches:
        n = min(a.check_every, a.launches - done)
        for i in range(n):
            if i % 8 == 0:
                ops.attention_fwd(qkv, B, T, H, scale, out=out, lse=lse)
            ops.attention_bwd(qkv, out, dout, lse, B, T, H, scale, dqkv=dqkv, delta=delta)
        torch.cuda.synchronize()
        done += n
        if not torch.equal(dqkv, ref_dqkv) or not torch.equal(out, ref_out):
            print(f"MISMATCH after {done} launches", flush=True)
            sys.exit(2)
    ctas = done * B * H
    print(f"{a.model} T={T}: {done} backward launches ({ctas / 1e6:.1f} M CTAs) + {done // 8} forward launches, all bit-identical, "
          f"{time.time() - t0:.1f} s", flush=True)


if __name__ == "__main__":
    main()
