"""Measures the GPU-side input pipeline (csrc/ingest.cu) on one B200: CUDA-event time per stage at batch 256, achieved HBM
GB/s against the algorithmic bytes of each stage, and the CPU restatement (oracle/ingest_oracle.py: the reference's per-image
numpy / cv2-style path) timed beside it on a bounded sample.   python tools/ingest_bench.py [--raw 512] [--batch 256]"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from thyroid_vit_cnn_comparison_b200 import ingest as ING, ops  # noqa: E402


def timed(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ms = []
    for _ in range(iters):
        flush.zero_()                                   # evict L2 (126 MB) between iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--raw", type=int, default=512)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--size", type=int, default=224)
    a = ap.parse_args()
    B, R, S = a.batch, a.raw, a.size
    rng = np.random.default_rng(0)
    raw_np = rng.integers(0, 65536, (B, R, R)).astype(np.uint16)
    raw = torch.from_numpy(raw_np.view(np.int16)).view(torch.uint16).cuda()
    gray = torch.empty(B, S, S, device="cuda")
    bounds = torch.empty(B, 2, device="cuda")
    out = torch.empty(B, 3, S, S, device="cuda")
    ing = ING.TileIngest(S, 3, percentiles=(1, 99))
    stages = {
        # bilinear taps touch ~all source rows when down-scaling by <= 2x: count the whole raw tile once
        "resize_u16": (lambda: ops.resize_u16(raw, S, S, out=gray), B * (R * R * 2 + S * S * 4)),
        "percentile_bounds": (lambda: ops.percentile_bounds(gray, 0.01, 0.99, out=bounds), B * S * S * 4),
        "finish_tiles": (lambda: ops.finish_tiles(gray, 3, bounds=bounds, mean=ING.IMAGENET_MEAN, std=ING.IMAGENET_STD, out=out),
                         B * (S * S * 4 + 3 * S * S * 4)),
        "whole pipeline": (lambda: ing(raw), B * (R * R * 2 + 3 * S * S * 4)),
    }
    res = {}
    for name, (fn, nbytes) in stages.items():
        ms = timed(fn)
        res[name] = {"ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1), "GB/s": round(nbytes / ms / 1e6, 1),
                     "images/s": round(B / ms * 1e3)}
    # CPU restatement on a bounded sample (per-image, as the reference's Dataset / AdaptiveNormalization do it)
    from oracle import ingest_oracle as IO
    n = 16
    t0 = time.perf_counter()
    x = torch.stack([IO.preprocess_image(raw_np[i], S) for i in range(n)])
    x = IO.to_channels_and_normalize(IO.adaptive_normalization(x), 3, ING.IMAGENET_MEAN, ING.IMAGENET_STD)
    dt = time.perf_counter() - t0
    res["cpu_port"] = {"images/s": round(n / dt, 1), "sample": f"{n} tiles {R}x{R} -> {S}x{S}, 1 process"}
    print(json.dumps({"batch": B, "raw": R, "size": S, "stages": res}, indent=1))


if __name__ == "__main__":
    main()
