"""SURVEY 8(f)-4 -- PTQ calibration with the forwards and the observer reductions on the B200.

The reference calibrates on the CPU: ``QuantizationEngine._calibrate(model, loader)`` (quantization/engines.py:123-133)
and the inline loop of quantization/main.py:236-239 call the observer-instrumented module that ``prepare_fx`` returned
once per calibration batch, and every observer keeps what it needs of the tensor it sees.  ``calibrate`` below stands
where those loops stand:

    prepared = prepare_fx(work_model, qconfig_mapping, example_inputs)     # unchanged (main.py:232)
    ievm_b200.calibrate(prepared, calib_loader)                             # instead of main.py:236-239
    q_model = convert_fx(prepared)                                          # unchanged (main.py:242)

* the float forward runs on the FP16 engine built from the prepared graph itself (``netdesc.from_prepared``: fused
  conv+BN(+ReLU) modules as they are, residual adds left un-fused so that the conv output in front of each add exists);
* ``ievm_observe`` reduces every observed tensor of the batch to ``torch.aminmax`` on the device, without a host sync
  between batches (``csrc/observe.cuh``); the network input is observed on the un-rounded f32 batch, so the input
  quantisation parameters are exactly the reference's;
* after the last batch the per-batch pairs are replayed, in the reference's call order, into the very observer modules
  of ``prepared`` -- an observer of the min/max family updates its state from ``aminmax(x)`` alone, so feeding it the
  2-element tensor ``[min, max]`` is the same update -- and ``convert_fx`` then runs unchanged.

Supported observers: ``MinMaxObserver`` and ``MovingAverageMinMaxObserver`` (the qconfig of quantization/main.py:196-207).
``HistogramObserver`` (the default fbgemm qconfig used by quantization/engines.py:103) needs the full histogram of
every tensor and is not built: ``calibrate`` raises ``NotImplementedError`` for it rather than approximating.

Activations are computed in fp16 with fp32 accumulation, the reference's calibration forward in fp32: scales agree to
about 1e-3 relative (tests/test_gpu_parity.py::test_gpu_calibration_*), they are not bit-identical.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .engine import B200HalfResNet
from .netdesc import POINT_LOGITS, POINT_POOLED, NetSpec, from_prepared


def check_observers(prepared, plan) -> None:
    """Every activation observer of the plan must keep nothing but a function of per-batch (min, max)."""
    from torch.ao.quantization.observer import MinMaxObserver
    for name, _ in plan:
        obs = prepared.get_submodule(name)
        if not isinstance(obs, MinMaxObserver):          # MovingAverageMinMaxObserver subclasses MinMaxObserver
            raise NotImplementedError(
                f"{name} is a {type(obs).__name__}: GPU calibration covers the min/max observer family "
                "(quantization/main.py:196-207); histogram observers (default fbgemm qconfig, quantization/engines.py:103) "
                "need a device histogram pass that is not built")


def point_index(point, num_tensors: int) -> int:
    """Column of an observation point in the log of ``ievm_observer_read`` (include/ievm.h)."""
    if point == POINT_POOLED:
        return num_tensors
    if point == POINT_LOGITS:
        return num_tensors + 1
    return int(point)


def replay_observers(prepared, plan: List[Tuple[str, object]], stats: np.ndarray, num_tensors: int) -> None:
    """Feed the recorded per-batch (min, max) pairs to the observers of ``prepared`` in the order the reference's
    calibration forward calls them (quantization/engines.py:130-133): batch by batch, node by node."""
    if stats.ndim != 3 or stats.shape[2] != 2:
        raise ValueError(f"expected [records, points, 2], got {stats.shape}")
    cols = [point_index(pt, num_tensors) for _, pt in plan]
    observers = [prepared.get_submodule(name) for name, _ in plan]
    if np.isnan(stats[:, cols]).any():
        raise RuntimeError("an observation point recorded no value (NaN activations or an unobserved tensor)")
    with torch.no_grad():
        for rec in stats:
            for obs, c in zip(observers, cols):
                obs(torch.tensor([rec[c, 0], rec[c, 1]], dtype=torch.float32))


class CalibrationEngine(B200HalfResNet):
    """FP16 engine over the prepared graph with one buffer per tensor and the device observer log switched on."""

    def __init__(self, net: NetSpec, **kw):
        super().__init__(net, **kw)
        self.set_option("keep_tensors", 1)
        self.set_option("observe", 1)
        self.num_points = int(self._lib.ievm_observer_points(self._handle))
        self.num_tensors = self.num_points - 2

    def observe(self, x_f32: Optional[torch.Tensor] = None) -> None:
        """Append the (min, max) pairs of the last forward to the device log (no synchronisation).  ``x_f32`` is the
        un-rounded f32 CUDA batch the forward's f16 input was cast from; without it the f16 input is observed."""
        ptr = None
        if x_f32 is not None:
            if not x_f32.is_cuda or x_f32.dtype != torch.float32 or not x_f32.is_contiguous():
                raise ValueError("x_f32 must be a contiguous float32 CUDA tensor")
            ptr = x_f32.data_ptr()
        stream = torch.cuda.current_stream(self.device_index).cuda_stream
        _lib.check(self._lib.ievm_observe(self._handle, ptr, stream), "ievm_observe")

    def read_observations(self) -> np.ndarray:
        """[records, points, 2] float32 = (min, max) per ``observe`` call and observation point (synchronises)."""
        cap = 4096
        buf = np.empty((cap, self.num_points, 2), np.float32)
        n = self._lib.ievm_observer_read(self._handle, buf.ctypes.data, cap)
        if n < 0:
            _lib.check(n, "ievm_observer_read")
        return buf[:n].copy()

    def reset_observations(self) -> None:
        self.set_option("observe", 1)


def _images_of(batch):
    return batch[0] if isinstance(batch, (tuple, list)) else batch


def calibrate(prepared, loader: Iterable, device: Optional[int] = None, max_batch: Optional[int] = None,
              return_stats: bool = False):
    """Drop-in for ``QuantizationEngine._calibrate(model, loader)`` (quantization/engines.py:123-133) and the loop at
    quantization/main.py:236-239: one observer update per calibration batch, forwards and reductions on the B200.
    ``loader`` yields ``(images, labels)`` pairs or image tensors (f32 NCHW, CPU or CUDA).  Afterwards ``prepared`` is
    in the state the reference's loop leaves it in (up to fp16 activation rounding) and ``convert_fx(prepared)``
    follows as in the reference."""
    prepared.eval()
    net, plan = from_prepared(prepared)
    check_observers(prepared, plan)
    dev = torch.cuda.current_device() if device is None else int(device)
    eng = None
    try:
        for batch in loader:
            images = _images_of(batch)
            if images.dim() != 4:
                raise ValueError(f"expected [N,C,H,W] images, got {tuple(images.shape)}")
            if eng is None:
                if tuple(images.shape[2:]) != (net.in_h, net.in_w):
                    net, plan = from_prepared(prepared, in_hw=tuple(images.shape[2:]))
                eng = CalibrationEngine(net, device=dev, max_batch=int(max_batch or images.shape[0]))
            if images.shape[0] > eng.max_batch:
                raise ValueError(f"calibration batch of {images.shape[0]} exceeds max_batch={eng.max_batch}; an observer "
                                 "update is per batch, so batches are not split -- pass max_batch")
            if images.shape[0] == 0:
                continue
            x32 = images.to(device=f"cuda:{dev}", dtype=torch.float32, non_blocking=True).contiguous()
            x16 = x32.half()
            logits = eng(x16)
            eng.observe(x32)
            del logits
        if eng is None:
            raise ValueError("empty calibration loader")
        stats = eng.read_observations()
        replay_observers(prepared, plan, stats, eng.num_tensors)
    finally:
        if eng is not None:
            eng.close()
    return stats if return_stats else None
