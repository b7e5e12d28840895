"""Stub: src/data/dataset.py imports tifffile at module top; only file loading uses it (not the arithmetic the oracle pins)."""


def imread(*a, **k):
    raise RuntimeError("tifffile stub: image file I/O is outside the oracle")
