"""Per-kernel CUDA time of one FrozenDenseNet forward (batch 256, bf16) under torch.profiler: which passes are left."""
import collections, sys
from pathlib import Path
import torch, torchvision
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import thyroid_vit_cnn_comparison_b200  # noqa
from thyroid_vit_cnn_comparison_b200 import teacher as T
torch.backends.cudnn.benchmark = True
m = torchvision.models.densenet169(weights=None, num_classes=2).cuda().eval().to(torch.bfloat16).to(memory_format=torch.channels_last)
x = torch.rand(256, 3, 224, 224, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
fast = T.FrozenDenseNet(m, dtype=torch.bfloat16)
with torch.no_grad():
    for _ in range(3):
        fast(x)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        fast(x)
        torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        n = ev.name
        key = n[:70]
        agg[key][0] += 1
        agg[key][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in agg.values())
print(f"total kernel time {tot / 1e3:.2f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{v[1] / 1e3:7.2f} ms  {v[0]:4d}x  {k}")
durs = [(ev.time_range.start, ev.device_time if hasattr(ev, "device_time") else ev.cuda_time) for ev in prof.events()
        if ev.device_type == torch.autograd.DeviceType.CUDA and "dense_bottleneck" in ev.name]
durs.sort()
print("bottleneck launches in order (us):", " ".join(f"{d:.0f}" for _, d in durs))
