"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the statistics the reference's PTQ observers take of a tensor.

Not product code (only tests/ import it).  The reference's calibration (quantization/engines.py:123-133,
quantization/main.py:236-239) calls torch.ao observers, un-vendored PyTorch code (requirements.txt:1, this image:
torch 2.11.0); what they read from a tensor is

  torch.aminmax(x)                                  every observer
  torch.histc(x, bins=2048, min=lo, max=hi)         HistogramObserver.forward, [lo, hi] = its running range

``histc_counts`` restates the bin assignment of ATen's CPU kernel (aten/src/ATen/native/cpu/HistogramKernel.cpp,
``histogramdd_cpu_contiguous`` with linear interpolation; outer edges from ``histc_select_outer_bin_edges``) in float32:

  lo == hi  ->  lo -= 1, hi += 1
  pos = int64(((x - lo) * bins) / (hi - lo));  pos == bins -> bins - 1

PARITY PIN: tests/test_calibration.py::test_histc_restatement_equals_torch_histc compares it with ``torch.histc`` itself
on seeded fp32 / fp16-valued / post-ReLU data, including ranges wider than the data and the degenerate lo == hi case.
The CUDA kernel (csrc/observe.cuh: observe_hist_multi_kernel) implements exactly this formula.
"""
import numpy as np


def histc_counts(x: np.ndarray, lo: float, hi: float, bins: int = 2048) -> np.ndarray:
    x = np.asarray(x, np.float32).ravel()
    lo, hi = np.float32(lo), np.float32(hi)
    if lo == hi:
        lo, hi = np.float32(lo - np.float32(1)), np.float32(hi + np.float32(1))
    x = x[(x >= lo) & (x <= hi)]                         # elements outside the range are ignored
    rng = np.float32(hi - lo)
    pos = ((x - lo).astype(np.float32) * np.float32(bins)).astype(np.float32) / rng
    pos = pos.astype(np.float32).astype(np.int64)
    pos[pos == bins] = bins - 1
    return np.bincount(pos, minlength=bins).astype(np.int64)
