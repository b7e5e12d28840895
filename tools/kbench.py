"""Per-kernel micro-benchmark at the bench shapes (DeiT-tiny / ViT-B, batch 256): CUDA-event timing of every
libvitk kernel class in isolation, L2 flushed between launches.  Also the short, deterministic command that
`ncu --set full` is pointed at (tools/kbench.py --model deit_tiny --only gemm_gelu --iters 2).

    python tools/kbench.py [--model deit_tiny|vit_base] [--only name[,name...]] [--iters 10] [--no-flush]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import thyroid_vit_cnn_comparison_b200  # noqa: E402,F401
from thyroid_vit_cnn_comparison_b200 import _lib, ops  # noqa: E402

DEV = "cuda"
F16 = torch.float16


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="deit_tiny")
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--json", default="")
    ap.add_argument("--warm", type=int, default=2, help="untimed launches per case (0 under ncu --set full: one launch per case)")
    a = ap.parse_args()
    B = a.batch
    if a.model == "deit_tiny":
        T, D, H, HID = 198, 192, 3, 768
    else:
        T, D, H, HID = 197, 768, 12, 3072
    M = B * T
    g = torch.Generator(device="cpu").manual_seed(0)
    r16 = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(DEV).to(F16)
    r32 = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(DEV)
    x16, x32, x32b = r16(M, D), r32(M, D), r32(M, D)
    h16, h16b = r16(M, HID), r16(M, HID)
    qkv, dqkv = r16(M, 3 * D), torch.empty(M, 3 * D, dtype=F16, device=DEV)
    w_qkv, w_proj, w_fc1, w_fc2 = r16(3 * D, D, sc=.05), r16(D, D, sc=.05), r16(HID, D, sc=.05), r16(D, HID, sc=.05)
    b_d, b_3d, b_h = r32(D), r32(3 * D), r32(HID)
    o16, o16b = torch.empty(M, D, dtype=F16, device=DEV), torch.empty(M, D, dtype=F16, device=DEV)
    o32 = torch.empty(M, D, device=DEV)
    oh, oh2 = torch.empty(M, HID, dtype=F16, device=DEV), torch.empty(M, HID, dtype=F16, device=DEV)
    o3 = torch.empty(M, 3 * D, dtype=F16, device=DEV)
    gw_fc1, gw_fc2 = torch.zeros(HID, D, device=DEV), torch.zeros(D, HID, device=DEV)
    gw_qkv, gw_proj = torch.zeros(3 * D, D, device=DEV), torch.zeros(D, D, device=DEV)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    gam, bet = r32(D), r32(D)
    dgam, dbet, dcs = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    lse, delta = torch.empty(B, H, T, device=DEV), torch.empty(B, H, T, device=DEV)
    cs_h, cs_3d = torch.zeros(HID, device=DEV), torch.zeros(3 * D, device=DEV)
    one = torch.ones(1, device=DEV)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    scale = 64 ** -0.5

    def splitk(m_out, n_out, k):
        bn = 256 if n_out % 256 == 0 else 192 if n_out % 192 == 0 else 128 if n_out % 128 == 0 else 64
        tiles = ((m_out + 127) // 128) * ((n_out + bn - 1) // bn)
        nkb = (k + 63) // 64
        return max(1, min(max(1, nkb // 2), (2 * sms + tiles - 1) // tiles))

    E = M * D * 2  # bytes of one 16-bit [M, D] tensor
    R = HID // D
    ops.layernorm_fwd(x32, gam, bet, y=o16, mean=mean, rstd=rstd)
    ops.attention_fwd(qkv, B, T, H, scale, out=o16b, lse=lse)
    cases = {
        # name: (fn, algorithmic bytes, flops)
        "ln_fwd": (lambda: ops.layernorm_fwd(x32, gam, bet, y=o16, mean=mean, rstd=rstd), 3 * E, 0),
        "ln_bwd": (lambda: ops.layernorm_bwd(x16, x32, mean, rstd, gam, dgam, dbet, dres=x32b, dx=o32, dx16=o16, dcolsum=dcs,
                                             unscale=one), 8 * E, 0),
        "gemm_qkv": (lambda: ops.gemm(x16, w_qkv, M, 3 * D, D, out=o3, bias=b_3d), 4 * E, 2 * M * 3 * D * D),
        "gemm_proj": (lambda: ops.gemm(x16, w_proj, M, D, D, out=o32, bias=b_d, residual=x32), 5 * E, 2 * M * D * D),
        "gemm_gelu": (lambda: ops.gemm(x16, w_fc1, M, HID, D, out=oh, out2=oh2, bias=b_h, epilogue=_lib.EPI_GELU),
                      (1 + 2 * R) * E, 2 * M * HID * D),
        "gemm_fc2": (lambda: ops.gemm(h16, w_fc2, M, D, HID, out=o32, bias=b_d, residual=x32), (R + 4) * E, 2 * M * HID * D),
        "dgrad_dgelu": (lambda: ops.gemm(x16, w_fc2, M, HID, D, b_mn=True, out=oh, aux=h16, epilogue=_lib.EPI_DGELU),
                        (1 + 2 * R) * E, 2 * M * HID * D),
        "dgrad_fc1": (lambda: ops.gemm(h16, w_fc1, M, D, HID, b_mn=True, out=o16), (R + 1) * E, 2 * M * HID * D),
        "dgrad_qkv": (lambda: ops.gemm(qkv, w_qkv, M, D, 3 * D, b_mn=True, out=o16), 4 * E, 2 * M * 3 * D * D),
        "dgrad_proj": (lambda: ops.gemm(x16, w_proj, M, D, D, b_mn=True, out=o16), 2 * E, 2 * M * D * D),
        "wgrad_fc2": (lambda: ops.gemm(x16, h16, D, HID, M, a_mn=True, b_mn=True, out=gw_fc2, epilogue=_lib.EPI_ATOMIC_ADD,
                                       split_k=0, alpha_dev=one), (R + 1) * E, 2 * M * HID * D),
        "wgrad_fc1": (lambda: ops.gemm(h16, x16, HID, D, M, a_mn=True, b_mn=True, out=gw_fc1, epilogue=_lib.EPI_ATOMIC_ADD,
                                       split_k=0, alpha_dev=one), (R + 1) * E, 2 * M * HID * D),
        "wgrad_qkv": (lambda: ops.gemm(qkv, x16, 3 * D, D, M, a_mn=True, b_mn=True, out=gw_qkv, epilogue=_lib.EPI_ATOMIC_ADD,
                                       split_k=0, alpha_dev=one), 4 * E, 2 * M * 3 * D * D),
        "wgrad_proj": (lambda: ops.gemm(x16, o16b, D, D, M, a_mn=True, b_mn=True, out=gw_proj, epilogue=_lib.EPI_ATOMIC_ADD,
                                        split_k=0, alpha_dev=one), 2 * E, 2 * M * D * D),
        "attn_fwd": (lambda: ops.attention_fwd(qkv, B, T, H, scale, out=o16b, lse=lse), 4 * E, 4 * B * H * T * T * 64),
        "attn_bwd": (lambda: ops.attention_bwd(qkv, o16b, x16, lse, B, T, H, scale, dqkv=dqkv, delta=delta), 9 * E,
                     10 * B * H * T * T * 64),
        "colsum_h": (lambda: ops.colsum16(h16, cs_h, unscale=one), R * E, 0),
        "colsum_3d": (lambda: ops.colsum16(qkv, cs_3d, unscale=one), 3 * E, 0),
    }
    # round-2 kernels around the encoder (input tiles, eval attention maps + rollout, the frozen teacher's bottleneck GEMM)
    want = set(s for s in a.only.split(",") if s)
    if a.model == "deit_tiny" and (not want or want & {"tiles_to_patches", "attn_probs", "rollout_row", "bottleneck_b1", "bottleneck_b3", "stem_conv7"}):
        BF = torch.bfloat16
        tiles = torch.randint(0, 65536, (B, 224, 224), dtype=torch.int32).to(torch.uint16).to(DEV)
        patches = torch.empty(B * 196, 768, dtype=F16, device=DEV)
        cases["tiles_to_patches"] = (lambda: ops.tiles_to_patches(tiles, 3, 16, out=patches), tiles.numel() * 2 + patches.numel() * 2, 0)
        probs = torch.empty(B, H, T, T, device=DEV)
        cases["attn_probs"] = (lambda: ops.attention_fwd(qkv, B, T, H, scale, out=o16b, lse=lse, probs=probs), 4 * E + probs.numel() * 4,
                               6 * B * H * T * T * 64)
        if not want or "rollout_row" in want:
            maps = torch.rand(B, 12, H, T, T, device=DEV)       # image-major, as ensemble.EnsembleInference keeps them
            cases["rollout_row"] = (lambda: ops.attention_rollout_row(maps, 0, "mean", image_major=True), maps.numel() * 4, 0)
        for nm, (P_, Ct_, C_) in {"bottleneck_b1": (B * 56 * 56, 256, 128), "bottleneck_b3": (B * 14 * 14, 1280, 768)}.items():
            if want and nm not in want:
                continue
            xb = torch.randn(P_, Ct_, generator=g).to(DEV).to(BF)
            wb = (torch.randn(128, C_, generator=g) * 0.05).to(DEV).to(BF)
            sb_, hb_, bb_ = torch.rand(C_, device=DEV) + 0.5, torch.randn(C_, device=DEV) * 0.1, torch.randn(128, device=DEV) * 0.1
            ob_ = torch.empty(P_, 128, dtype=BF, device=DEV)
            cases[nm] = ((lambda xb=xb, C_=C_, sb_=sb_, hb_=hb_, wb=wb, bb_=bb_, ob_=ob_: ops.dense_bottleneck(xb, C_, sb_, hb_, wb, bb_, out=ob_)),
                         P_ * C_ * 2 + P_ * 128 * 2 + 128 * C_ * 2, 2 * P_ * C_ * 128)
        if not want or "stem_conv7" in want:
            xs = torch.randn(B, 224, 224, 3, generator=g).to(DEV).to(BF)
            ws = ops.stem_conv7_weights(torch.randn(64, 3, 7, 7, generator=g) * 0.1, BF).to(DEV)
            bs_ = torch.randn(64, device=DEV)
            os_ = torch.empty(B, 112, 112, 64, dtype=BF, device=DEV)
            cases["stem_conv7"] = (lambda: ops.stem_conv7(xs, ws, bs_, out=os_), xs.numel() * 2 + os_.numel() * 2, 2 * B * 112 * 112 * 64 * 147)
    only = [s for s in a.only.split(",") if s]
    flush = None if a.no_flush else torch.empty(256 * 1024 * 1024 // 4, device=DEV)
    res = {}
    for name, (fn, nbytes, flops) in cases.items():
        if only and name not in only:
            continue
        for _ in range(a.warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.iters):
            if flush is not None:
                flush.zero_()
            torch.cuda._sleep(400000)  # ~0.2 ms of GPU idle spinning: the host enqueues the launch behind it, so the
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()                # events bracket device time only (no host launch latency inside)
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        res[name] = {"us": ms * 1e3, "gbs": nbytes / ms / 1e6, "tflops": flops / ms / 1e9, "ideal_us_hbm": nbytes / 6544.3e9 * 1e6}
        print(f"{a.model:9s} {name:12s} {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.0f} GB/s  {flops / ms / 1e9:7.1f} TF  "
              f"(HBM floor {nbytes / 6544.3e9 * 1e6:6.1f} us)", flush=True)
    if a.json:
        Path(a.json).write_text(json.dumps({"model": a.model, "batch": B, "kernels": res}, indent=1))


if __name__ == "__main__":
    main()
