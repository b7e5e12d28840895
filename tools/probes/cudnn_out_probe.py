"""Can cuDNN write a convolution's output straight into a channel slice of an NHWC concatenation buffer
(aten::cudnn_convolution.out with a strided `out`)?  Checks the values, that the other channels stay untouched, and the time
against conv2d + the channel-slice write."""
import torch
import torch.nn.functional as F

torch.backends.cudnn.benchmark = True
dev, dt = "cuda", torch.bfloat16
for (B, HW, ct, c) in ((256, 56, 256, 64), (256, 28, 512, 128), (256, 14, 1280, 256), (256, 7, 1664, 640)):
    g = torch.Generator().manual_seed(1)
    t = torch.randn(B, 128, HW, HW, generator=g).to(dev, dt).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(32, 128, 3, 3, generator=g) * 0.05).to(dev, dt).contiguous(memory_format=torch.channels_last)
    buf = torch.full((B, HW, HW, ct), 7.0, dtype=dt, device=dev)
    ref = F.conv2d(t, w, None, 1, 1)
    view = buf[..., c:c + 32].permute(0, 3, 1, 2)          # [B,32,H,W], strides (H*W*ct, 1, W*ct, ct)
    try:
        torch.ops.aten.cudnn_convolution.out(t, w, [1, 1], [1, 1], [1, 1], 1, True, False, True, out=view)
    except Exception as e:          # noqa: BLE001
        print("FAILED", (B, HW, ct, c), repr(e)[:300])
        continue
    torch.cuda.synchronize()
    ok = torch.equal(buf[..., c:c + 32], ref.permute(0, 2, 3, 1))
    close = (buf[..., c:c + 32].float() - ref.permute(0, 2, 3, 1).float()).abs().max().item()
    untouched = bool((buf[..., :c] == 7).all() and (buf[..., c + 32:] == 7).all())

    def timeit(fn):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / 20 * 1e3

    t_direct = timeit(lambda: torch.ops.aten.cudnn_convolution.out(t, w, [1, 1], [1, 1], [1, 1], 1, True, False, True, out=view))
    t_two = timeit(lambda: buf[..., c:c + 32].copy_(F.conv2d(t, w, None, 1, 1).permute(0, 2, 3, 1)))
    t_conv = timeit(lambda: F.conv2d(t, w, None, 1, 1))
    print(f"HW={HW} ct={ct}: equal={ok} maxdiff={close:.3g} untouched={untouched} direct {t_direct:.1f} us, conv {t_conv:.1f} us, conv+aten copy {t_two:.1f} us")
