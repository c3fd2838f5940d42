"""Shared helpers for the test tiers."""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CACHE = {}


def cached_quantized(widths):
    """Converted INT8 module for `widths` (PTQ takes a few seconds; share it across tests)."""
    from oracle import model_factory as mf
    key = ("q", tuple(widths))
    if key not in _CACHE:
        _CACHE[key] = mf.static_quantize_fbgemm(mf.make_student(widths))
    return _CACHE[key]
