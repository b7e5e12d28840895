"""GEMM time vs number of tile rounds (fixed overhead vs per-tile cost).  VITK_LIB / VITK_GEMM_KNOBS honoured."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import thyroid_vit_cnn_comparison_b200  # noqa
from thyroid_vit_cnn_comparison_b200 import _lib, ops
F16 = torch.float16
N, K = int(sys.argv[1]), int(sys.argv[2])
for rounds in (1, 2, 4, 8, 16):
    bn = 256 if N % 256 == 0 else 192 if N % 192 == 0 else 128 if N % 128 == 0 else 64
    ntiles = N // bn
    M = 128 * 148 * rounds // ntiles
    A = torch.randn(M, K, device="cuda").to(F16)
    W = (torch.randn(N, K, device="cuda") * .05).to(F16)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, dtype=F16, device="cuda")
    fn = lambda: ops.gemm(A, W, M, N, K, out=out, bias=bias)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(8):
        torch.cuda._sleep(400000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"N={N} K={K} M={M} tiles/SM={rounds} min {ts[0]:.1f} med {ts[4]:.1f} us")
