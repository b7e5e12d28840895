"""B200-native ViT/DeiT training path: drop-in for gogolB/thyroid-vit-cnn-comparison.

Package directory is `thyroid-vit-cnn-comparison_b200/` (repo naming contract); import it as
`thyroid_vit_cnn_comparison_b200` via the shim module at the repo root.

Everything numeric runs in libvitk.so (hand-written sm_100a CUDA behind the C-ABI of
include/vitk.h).  Importing the package does not need a GPU; running any op does.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
