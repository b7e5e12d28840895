"""Where the frozen DenseNet169 teacher forward spends its time (distillation step): per-kernel-class CUDA
time of one bf16 channels_last forward at batch 256 under torch.profiler.  The teacher is library code (cuDNN / ATen); this is
the evidence for which of its passes a fused kernel would remove.

    python tools/teacher_profile.py [--batch 256]
"""
import argparse
import collections

import sys
from pathlib import Path

import torch
import torchvision

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    a = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    m = torchvision.models.densenet169(weights=None, num_classes=2).cuda().eval().to(torch.bfloat16).to(memory_format=torch.channels_last)
    x = torch.rand(a.batch, 3, 224, 224, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        for _ in range(3):
            m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            m(x)
        e1.record()
        torch.cuda.synchronize()
        print(f"eager forward: {e0.elapsed_time(e1) / 5:.2f} ms / batch {a.batch}")
        # the same network through the frozen-teacher executor (teacher.py: block buffer, fused BatchNorm + ReLU, folded convs)
        import thyroid_vit_cnn_comparison_b200  # noqa: F401
        from thyroid_vit_cnn_comparison_b200 import teacher as T
        fast = T.FrozenDenseNet(m, dtype=torch.bfloat16)
        for _ in range(3):
            fast(x)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            fast(x)
        e1.record()
        torch.cuda.synchronize()
        print(f"FrozenDenseNet forward: {e0.elapsed_time(e1) / 5:.2f} ms / batch {a.batch} (fused conv+bias+relu: {fast._fused_conv_relu})")
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            m(x)
            torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            n = ev.name
            key = ("conv/gemm" if any(s in n for s in ("gemm", "conv", "cutlass", "xmma", "sm90", "sm100", "nvjet", "cudnn")) else
                   "batch_norm" if "batch_norm" in n or "bn_" in n else
                   "cat/copy" if "Cat" in n or "copy" in n.lower() else
                   "relu/elementwise" if "elementwise" in n else
                   "pool" if "pool" in n.lower() else n[:60])
            agg[key][0] += 1
            agg[key][1] += ev.device_time
    tot = sum(v[1] for v in agg.values())
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:60s} n={c:4d} {t / 1e3:8.3f} ms  {t / tot:6.1%}")
    print(f"total device time {tot / 1e3:.2f} ms")


if __name__ == "__main__":
    main()
