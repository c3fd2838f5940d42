"""TEST INFRASTRUCTURE (oracle) -- restatement of Pillow's 8-bit bilinear resize, the first step of the reference's
dataset transform: ``T.Resize((224, 224))`` on a PIL image (quantization/dataset.py:15) calls
``Image.resize(size, BILINEAR)`` -> ``ImagingResample`` (Pillow ``src/libImaging/Resample.c``, not vendored in the
reference; pinned by the installed Pillow, checked in tests/test_oracle.py against ``PIL.Image.resize`` itself).

Algorithm (8 bits per channel): a horizontal pass then a vertical pass, each a per-output-pixel weighted sum of the
input pixels inside the filter support, with fixed-point coefficients of PRECISION_BITS = 22 bits, a rounding bias of
half an LSB, and a clip to [0, 255]; the intermediate image is rounded to 8 bits between the passes.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bilinear(x: float) -> float:
    x = -x if x < 0.0 else x
    return 1.0 - x if x < 1.0 else 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c:precompute_coeffs + normalize_coeffs_8bpc for the whole-image box and the bilinear filter.
    Returns (bounds int32 [out, 2] = (first input index, tap count), kk int32 [out, ksize])."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bilinear((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w)                               # C: running sum in the same left-to-right order
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray, axis: int) -> np.ndarray:
    src = np.moveaxis(img.astype(np.int64), axis, 0)          # resampled axis first
    out = np.empty((bounds.shape[0],) + src.shape[1:], np.int64)
    for xx in range(bounds.shape[0]):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(n):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """img: uint8 [H, W, C] (or [N, H, W, C]).  Horizontal pass first, then vertical, like ImagingResample."""
    h_axis, w_axis = img.ndim - 3, img.ndim - 2
    out = img
    if img.shape[w_axis] != out_w:
        out = _pass(out, *precompute_coeffs(img.shape[w_axis], out_w), axis=w_axis)
    if img.shape[h_axis] != out_h:
        out = _pass(out, *precompute_coeffs(img.shape[h_axis], out_h), axis=h_axis)
    return out
