"""Profiling build of libvitk (libvitk_dbg.so, -DVITK_GEMM_KNOBS): the GEMM kernel then honours the env var
VITK_GEMM_KNOBS (bit 1: no TMA stores, 2: no epilogue math/stores, 4: B operand loaded once per CTA, 8: A likewise) so that
the cost of each phase can be measured by elimination.  Use with VITK_LIB=thyroid-vit-cnn-comparison_b200/libvitk_dbg.so."""
import subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "thyroid-vit-cnn-comparison_b200"
sys.path.insert(0, str(PKG))
import build as B  # noqa: E402
out = PKG / "build_dbg"
out.mkdir(exist_ok=True)
objs = []
for src in B.SOURCES:
    obj = out / src.replace(".cu", ".o")
    flags = [f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + ["-DVITK_GEMM_KNOBS"]
    subprocess.run([B._nvcc(), *flags, "-c", str(B.CSRC / src), "-o", str(obj)], check=True)
    objs.append(str(obj))
subprocess.run([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(PKG / "libvitk_dbg.so"), *objs], check=True)
print(PKG / "libvitk_dbg.so")
