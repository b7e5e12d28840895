"""Timeline of CTA (0,0) of the tcgen05 attention backward (profiling build): SM-clock stamps of the MMA thread and of
math warp 0."""
import ctypes as C, os, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
os.environ["VITK_LIB"] = str(ROOT / "thyroid-vit-cnn-comparison_b200" / "libvitk_dbg.so")
sys.path.insert(0, str(ROOT))
import thyroid_vit_cnn_comparison_b200  # noqa
from thyroid_vit_cnn_comparison_b200 import _lib, ops
B, T, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
F16 = torch.float16
qkv = torch.randn(B, T, 3 * H * 64, device="cuda").to(F16)
out, lse = ops.attention_fwd(qkv, B, T, H, 0.125)
dout = torch.randn(B, T, H * 64, device="cuda").to(F16)
for _ in range(3):
    ops.attention_bwd(qkv, out, dout, lse, B, T, H, 0.125)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_longlong * 256)()
lib.vitk_debug_read_attn.argtypes = [C.c_void_p]
assert lib.vitk_debug_read_attn(buf) == 0
v = list(buf)
t0 = v[0]
print("MMA thread: start 0 | K,Q landed", v[1] - t0, "| V,dO landed", v[2] - t0)
NT = ((T + 127) // 128) * ((((T + 15) // 16) * 16 + 63) // 64)
print("t | mma: ld_ok p_ok issued | math: top s_ok ld_done m2_ok p_arrived [epi_done]")
for t in range(NT):
    m = [v[4 + 4 * t + i] - t0 for i in range(3)]
    w = [v[68 + 8 * t + i] - t0 for i in range(6)]
    print(f"{t:2d} | {m[0]:7d} {m[1]:7d} {m[2]:7d} | {w[0]:7d} {w[1]:7d} {w[2]:7d} {w[3]:7d} {w[4]:7d} {w[5] if w[5] > 0 else 0:7d}")
print("math: delta done", v[64] - t0, "| all done", v[65] - t0)

f = [v[200 + i] for i in range(8)]
f0 = f[0]
print("forward CTA (tile 0, head 0, image 100), cycles from the MMA thread's start:")
print("  Q,K landed", f[1] - f0, "| S ready (softmax starts)", f[3] - f0, "| max pass done", f[4] - f0, "| P written", f[5] - f0,
      "| PV issued", f[2] - f0, "| O ready", f[6] - f0, "| stored", f[7] - f0)
