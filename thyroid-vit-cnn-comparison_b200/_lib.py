"""ctypes binding of libvitk.so -- the ONLY compute backend of this package.

There is deliberately no fallback: if the shared library is missing and cannot be built, or a
kernel returns an error, a RuntimeError is raised.  Signatures mirror include/vitk.h one to one.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libvitk.so"

c_void_p, c_int, c_int64, c_float = C.c_void_p, C.c_int32, C.c_int64, C.c_float

ABI_VERSION = 25
DT_BF16, DT_FP32, DT_FP16 = 0, 1, 2
EPI_STORE, EPI_GELU, EPI_DGELU, EPI_ATOMIC_ADD, EPI_TOKENS = 0, 1, 2, 3, 4


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", c_void_p), ("B", c_void_p),
        ("lda", c_int64), ("ldb", c_int64),
        ("a_mn_major", c_int), ("b_mn_major", c_int),
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("split_k", c_int), ("epilogue", c_int), ("out_dtype", c_int),
        ("a_dtype", c_int), ("b_dtype", c_int), ("aux_dtype", c_int),
        ("alpha", c_float), ("alpha_dev", c_void_p),
        ("bias", c_void_p), ("residual", c_void_p), ("ldr", c_int64),
        ("out", c_void_p), ("ldo", c_int64),
        ("out2", c_void_p), ("ldo2", c_int64),
        ("aux", c_void_p), ("ldaux", c_int64),
        ("rows_per_img", c_int), ("tokens_per_img", c_int), ("prefix", c_int),
        ("pos", c_void_p),
        ("colsum_out", c_void_p),
        ("row_scale", c_void_p),
        ("drop_seed", c_void_p), ("drop_p", c_float), ("drop_site", c_int),
    ]


class Dropout(C.Structure):
    """vitk_dropout: (device seed pointer, drop probability, site id)."""
    _fields_ = [("seed", c_void_p), ("p", c_float), ("site", c_int)]


# name -> (restype, argtypes); every symbol include/vitk.h declares
SIGNATURES = {
    "vitk_abi_version": (c_int, []),
    "vitk_last_error": (C.c_char_p, []),
    "vitk_launch_count": (c_int64, []),
    "vitk_reset_launch_count": (None, []),
    "vitk_gemm": (c_int, [C.POINTER(GemmArgs), c_void_p]),
    "vitk_layernorm_fwd": (c_int, [c_void_p] * 4 + [c_int] + [c_void_p] * 2 + [c_int64, c_int, c_float, c_void_p]),
    "vitk_layernorm_bwd": (c_int, [c_void_p, c_int] + [c_void_p] * 7 + [c_int] + [c_void_p] * 6 + [c_int64, c_int, c_void_p]),
    "vitk_attention_fwd": (c_int, [c_void_p] * 2 + [c_int] + [c_void_p] * 2 + [c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "vitk_attention_bwd": (c_int, [c_void_p] * 6 + [c_int] * 4 + [c_float, c_int, c_void_p]),
    "vitk_attention_dropout_fwd": (c_int, [c_void_p] * 2 + [c_int] + [c_void_p] + [c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "vitk_attention_dropout_bwd": (c_int, [c_void_p] * 6 + [c_int] * 4 + [c_float, c_void_p, c_void_p]),
    "vitk_patchify": (c_int, [c_void_p, c_void_p] + [c_int] * 6 + [c_void_p]),
    "vitk_patchify_hwc": (c_int, [c_void_p, c_void_p] + [c_int] * 6 + [c_void_p]),
    "vitk_prefix_tokens_fwd": (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_void_p, c_void_p]),
    "vitk_tokens_bwd": (c_int, [c_void_p] * 5 + [c_int] + [c_void_p] * 2 + [c_int] * 4 + [c_void_p, c_void_p]),
    "vitk_gather_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, C.c_int64, c_void_p]),
    "vitk_expand_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, C.c_int64, c_void_p]),
    "vitk_head_fwd": (c_int, [c_void_p] * 12 + [c_int] * 5 + [c_float, c_void_p]),
    "vitk_head_bwd": (c_int, [c_void_p] * 10 + [c_int] + [c_void_p] * 10 + [c_int] * 5 + [c_void_p]),
    "vitk_pool_norm_fwd": (c_int, [c_void_p] * 6 + [c_int] * 5 + [c_float, c_void_p]),
    "vitk_pool_norm_bwd": (c_int, [c_void_p] * 7 + [c_int] + [c_void_p] * 6 + [c_int] * 5 + [c_void_p]),
    "vitk_dense_fwd": (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_void_p]),
    "vitk_dense_bwd": (c_int, [c_void_p] * 8 + [c_int] * 4 + [c_void_p]),
    "vitk_affine_relu_nhwc": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "vitk_im2col_rows": (c_int, [c_void_p, c_void_p] + [c_int] * 7 + [c_int64, c_void_p]),
    "vitk_stem_conv7": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_void_p]),
    "vitk_dense_bottleneck": (c_int, [c_void_p, c_int64] + [c_void_p] * 5 + [c_int64, c_int, c_int, c_void_p]),
    "vitk_pool_nhwc": (c_int, [c_void_p, c_void_p, c_int64] + [c_int] * 9 + [c_void_p]),
    "vitk_dropout_mask": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "vitk_droppath_scale": (c_int, [c_void_p] * 3 + [c_int] * 3 + [c_void_p]),
    "vitk_resize_u16": (c_int, [c_void_p, c_void_p] + [c_int] * 5 + [c_void_p]),
    "vitk_percentile_bounds": (c_int, [c_void_p, c_int, c_int64, c_float, c_float, c_void_p, c_void_p]),
    "vitk_finish_tiles": (c_int, [c_void_p] * 3 + [c_int] * 4 + [c_void_p] * 3 + [c_int, c_float] + [c_int] * 4 + [c_void_p]),
    "vitk_tiles_to_patches": (c_int, [c_void_p, c_int] + [c_void_p] * 4 + [c_int] * 6 + [c_void_p]),
    "vitk_metrics_update": (c_int, [c_void_p, c_void_p, c_int, c_int] + [c_void_p] * 4 + [c_int64, c_void_p]),
    "vitk_binary_auroc": (c_int, [c_void_p] * 3 + [c_int64] + [c_void_p] * 3),
    "vitk_loss_fwd_bwd": (c_int, [c_void_p] * 7 + [c_int, c_int, c_int] + [c_float] * 5 + [c_void_p]),
    "vitk_sqnorm_scratch_floats": (c_int, []),
    "vitk_grad_sqnorm": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "vitk_adamw_step": (c_int, [c_void_p] * 10 + [c_int, c_void_p, c_void_p] + [c_float] * 4 + [c_int, c_void_p]),
    "vitk_amp_update": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "vitk_cast_f32_to_16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "vitk_colsum16": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "vitk_ensemble_probs": (c_int, [c_void_p] * 4 + [c_int] * 3 + [c_void_p]),
    "vitk_attention_rollout_row": (c_int, [c_void_p] * 2 + [c_int64] * 2 + [c_int] * 6 + [c_void_p]),
    "vitk_attention_probs": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_float, c_void_p]),
    "vitk_cls_attention_heatmap": (c_int, [c_void_p] * 2 + [c_int64] * 2 + [c_int] * 6 + [c_void_p]),
    "vitk_attention_rollout": (c_int, [c_void_p] * 3 + [c_int] * 5 + [c_void_p]),
}

_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load libvitk.so (building it with nvcc first if it is absent). Raises on any failure."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if not build_if_missing or os.environ.get("VITK_NO_BUILD"):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python {_HERE / 'build.py'}` "
                               "(there is no CPU / eager fallback)")
        import importlib.util
        spec = importlib.util.spec_from_file_location("_vitk_build", _HERE / "build.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    # VITK_LIB: load another build of the same C-ABI (e.g. the profiling build of tools/build_dbg.py)
    lib = C.CDLL(os.environ.get("VITK_LIB") or str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header / library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.vitk_abi_version() != ABI_VERSION:
        raise RuntimeError("libvitk.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().vitk_last_error().decode(errors="replace")
        raise RuntimeError(f"libvitk {what} failed (status {rc}): {msg}")
