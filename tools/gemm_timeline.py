"""Per-tile timeline of CTA 0 of one GEMM (profiling build, VITK_GEMM_KNOBS bit 64): where the producer, the MMA thread
and epilogue warp 2 spend their cycles."""
import ctypes as C, os, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
os.environ["VITK_LIB"] = str(ROOT / "thyroid-vit-cnn-comparison_b200" / "libvitk_dbg.so")
os.environ["VITK_GEMM_KNOBS"] = str(64 | int(os.environ.get("EXTRA_KNOBS", "0")))
sys.path.insert(0, str(ROOT))
import thyroid_vit_cnn_comparison_b200  # noqa
from thyroid_vit_cnn_comparison_b200 import _lib, ops
F16 = torch.float16
M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mode = sys.argv[4] if len(sys.argv) > 4 else "store"
if mode != "wgrad":
    A = torch.randn(M, K, device="cuda").to(F16)
    W = (torch.randn(N, K, device="cuda") * .05).to(F16)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, dtype=F16, device="cuda")
    out2 = torch.empty(M, N, dtype=F16, device="cuda")
    o32, r32 = torch.empty(M, N, device="cuda"), torch.randn(M, N, device="cuda")
if mode == "store":
    fn = lambda: ops.gemm(A, W, M, N, K, out=out, bias=bias)
elif mode == "gelu":
    fn = lambda: ops.gemm(A, W, M, N, K, out=out, out2=out2, bias=bias, epilogue=_lib.EPI_GELU)
elif mode == "wgrad":   # dW[M=N_out, N=K_in] += dY^T X over K = tokens;  argv: N_out K_in tokens
    dY = torch.randn(K, M, device="cuda").to(F16)
    X = torch.randn(K, N, device="cuda").to(F16)
    gw = torch.zeros(M, N, device="cuda")
    one = torch.ones(1, device="cuda")
    bn = 256 if N % 256 == 0 else 192 if N % 192 == 0 else 128 if N % 128 == 0 else 64
    tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
    nkb = (K + 63) // 64
    split = int(os.environ.get("SPLIT", "0")) or max(1, min(max(1, nkb // 2), (2 * 148 + tiles - 1) // tiles))
    print("split_k", split)
    fn = lambda: ops.gemm(dY, X, M, N, K, a_mn=True, b_mn=True, out=gw, epilogue=_lib.EPI_ATOMIC_ADD, split_k=split, alpha_dev=one)
else:
    fn = lambda: ops.gemm(A, W, M, N, K, out=o32, bias=bias, residual=r32)
for _ in range(3):
    fn()
torch.cuda.synchronize()
ts = []
for _ in range(8):   # event-timed duration of the same launch (operands hot in L2), for the cycles -> time conversion
    torch.cuda._sleep(400000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
print(f"event-timed launch (hot L2): min {ts[0]:.1f} med {ts[4]:.1f} us")
lib = _lib.load()
buf = (C.c_longlong * (3 * 64 * 4))()
lib.vitk_debug_read.argtypes = [C.c_void_p]
assert lib.vitk_debug_read(buf) == 0
v = list(buf)
g = lambda r, t, e: v[(r * 64 + t) * 4 + e]
t0 = g(1, 0, 0)
print(f"M={M} N={N} K={K} mode={mode}: cycles relative to the MMA thread's first stamp")
print("tile | prod:empty_ok | mma: start tempty_ok full0_ok issued | epi: start tfull_ok done | epi_busy mma_wait_tempty")
for t in range(12):
    if g(1, t, 0) == 0: break
    print(f"{t:4d} | {g(0,t,0)-t0:8d} | {g(1,t,0)-t0:8d} {g(1,t,1)-t0:8d} {g(1,t,2)-t0:8d} {g(1,t,3)-t0:8d} | {g(2,t,0)-t0:8d} {g(2,t,1)-t0:8d} {g(2,t,2)-t0:8d} | "
          f"{g(2,t,2)-g(2,t,1):6d} {g(1,t,1)-g(1,t,0):6d}")

buf2 = (C.c_longlong * (8 * 4 * 8))()
lib.vitk_debug_read2.argtypes = [C.c_void_p]
assert lib.vitk_debug_read2(buf2) == 0
w = list(buf2)
print("epilogue warp 2, per unit: tmem_ld | math+pack | wait_read | sts+fence | store_issue   (cycles)")
for t in range(2, 6):
    for u in range(4):
        e = [w[((t * 4 + u) * 8) + i] for i in range(6)]
        if e[0] == 0 or e[5] == 0: continue
        print(f"  tile {t} unit {u}: start {e[0]-t0:8d} | {e[1]-e[0]:5d} | {e[2]-e[1]:5d} | {e[3]-e[2]:5d} | {e[4]-e[3]:5d} | {e[5]-e[4]:5d}")

buf3 = (C.c_longlong * (256 * 4))()
lib.vitk_debug_read3.argtypes = [C.c_void_p]
assert lib.vitk_debug_read3(buf3) == 0
g3 = [[buf3[c * 4 + e] for e in range(4)] for c in range(256) if buf3[c * 4] != 0]
t_first = min(r[0] for r in g3)
t_last = max(r[3] for r in g3)
import statistics as st
print(f"wall clock over {len(g3)} CTAs (globaltimer ns): first entry -> last exit {t_last - t_first}")
for name, f in (("entry skew (entry - first entry)", lambda r: r[0] - t_first), ("set-up (entry -> set-up done)", lambda r: r[1] - r[0]),
                ("tile loop (set-up done -> last epilogue)", lambda r: r[2] - r[1]), ("teardown (last epilogue -> exit)", lambda r: r[3] - r[2]),
                ("exit slack (last exit - exit)", lambda r: t_last - r[3])):
    v = sorted(f(r) for r in g3)
    print(f"  {name:44s} min {v[0]:6d}  med {v[len(v)//2]:6d}  max {v[-1]:6d}")
