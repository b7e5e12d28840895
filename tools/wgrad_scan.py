"""Weight-gradient GEMM time against the K split (DeiT-tiny / ViT-B shapes, 50688 token rows), L2 flushed, median of 12.
Usage: python tools/wgrad_scan.py [tiny|base]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import thyroid_vit_cnn_comparison_b200  # noqa: F401,E402
from thyroid_vit_cnn_comparison_b200 import _lib, ops  # noqa: E402

F16 = torch.float16
which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
D = 192 if which == "tiny" else 768
M = 50688
shapes = {"qkv": (3 * D, D), "proj": (D, D), "fc1": (4 * D, D), "fc2": (D, 4 * D)}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
unscale = torch.ones(1, device="cuda")
for name, (n_out, k_in) in shapes.items():
    dy = (torch.randn(M, n_out, device="cuda") * 0.1).to(F16)
    x = (torch.randn(M, k_in, device="cuda") * 0.1).to(F16)
    gw = torch.zeros(n_out, k_in, device="cuda")
    gb = torch.zeros(n_out, device="cuda")
    line = []
    for sk in ([int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else (0, 8, 12, 16, 24, 32, 48, 64, 96)):
        try:
            fn = lambda: ops.gemm(dy, x, n_out, k_in, M, a_mn=True, b_mn=True, out=gw, epilogue=_lib.EPI_ATOMIC_ADD, split_k=sk,
                                  alpha_dev=unscale, colsum_out=gb if name in ("qkv", "fc1") else None)
            for _ in range(2):
                fn()
            ts = []
            for _ in range(12):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            ts.sort()
            line.append(f"sk={sk}:{ts[6]:.1f}")
        except Exception as e:      # noqa: BLE001
            line.append(f"sk={sk}:ERR")
    floor = (M * (n_out + k_in) * 2) / 6544.3e9 * 1e6
    print(f"{which} wgrad_{name} [{n_out}x{k_in}] HBM floor {floor:.1f} us | " + "  ".join(line), flush=True)
