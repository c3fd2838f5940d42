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

Supported observers: ``MinMaxObserver`` and ``MovingAverageMinMaxObserver`` (the qconfig of quantization/main.py:196-207)
need only the pairs.  ``HistogramObserver`` (the default fbgemm qconfig of ``QuantizationEngine.static_quantize``,
quantization/engines.py:103) additionally needs, per batch, ``torch.histc(x, 2048, min, max)`` over the observer's
running range: the device keeps that range per observer and bins every tensor with ATen's own float32 bin formula
(exact integer counts); the replay restates the ten lines of ``HistogramObserver.forward`` around the observer's own
``_combine_histograms``.  Any other observer class raises ``NotImplementedError``.

Activations are computed in fp16 with fp32 accumulation, the reference's calibration forward in fp32: observer ranges agree
to about 1e-3 relative (tests/test_zz_gpu_calibration.py, measured 8.7e-4), they are not bit-identical; given the same
tensors the device statistics and the replay are exact (tests/test_calibration.py, tests/test_zz_gpu_calibration.py).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .engine import B200HalfResNet
from .netdesc import POINT_LOGITS, POINT_POOLED, NetSpec, from_prepared


HIST_BINS = 2048          # kObsBins in csrc/observe.cuh == HistogramObserver's default


def observer_mode(prepared, plan) -> int:
    """1 when every activation observer of the plan keeps a function of per-batch (min, max) only (``MinMaxObserver``,
    ``MovingAverageMinMaxObserver``), 2 when ``HistogramObserver`` instances are among them (device histogram pass)."""
    from torch.ao.quantization.observer import HistogramObserver, MinMaxObserver
    mode = 1
    for name, _ in plan:
        obs = prepared.get_submodule(name)
        if isinstance(obs, HistogramObserver):
            if obs.bins != HIST_BINS:
                raise NotImplementedError(f"{name}: HistogramObserver with {obs.bins} bins (the device pass has {HIST_BINS})")
            mode = 2
        elif not isinstance(obs, MinMaxObserver):        # MovingAverageMinMaxObserver subclasses MinMaxObserver
            raise NotImplementedError(
                f"{name} is a {type(obs).__name__}: GPU calibration covers MinMaxObserver, MovingAverageMinMaxObserver "
                "(quantization/main.py:196-207) and HistogramObserver (default fbgemm qconfig, quantization/engines.py:103)")
    return mode


def check_observers(prepared, plan) -> None:
    observer_mode(prepared, plan)


def point_index(point, num_tensors: int) -> int:
    """Column of an observation point in the logs of ``ievm_observer_read*`` (include/ievm.h)."""
    if point == POINT_POOLED:
        return num_tensors
    if point == POINT_LOGITS:
        return num_tensors + 1
    return int(point)


def point_groups(prepared, plan, num_tensors: int) -> np.ndarray:
    """int32 [points]: the observer group of every observation point = the first point that reports to the same
    observer instance (prepare_fx shares one instance between max-pool / avg-pool / flatten and their input)."""
    groups = np.arange(num_tensors + 2, dtype=np.int32)
    first = {}
    for name, pt in plan:
        col = point_index(pt, num_tensors)
        groups[col] = first.setdefault(id(prepared.get_submodule(name)), col)
    return groups


def _replay_histogram(obs, x_min: torch.Tensor, x_max: torch.Tensor, counts: np.ndarray, used_range, where: str) -> None:
    """``HistogramObserver.forward`` (torch/ao/quantization/observer.py) with ``torch.aminmax(x)`` and
    ``torch.histc(x, bins, min, max)`` replaced by the values the device computed; everything else -- the running range,
    ``_combine_histograms`` with its up-scaling -- is the observer's own code on its own state."""
    import inspect
    # The replay calls the observer's private _combine_histograms; its signature changed between torch releases
    # (older ones: orig_hist, new_hist, upsample_rate, downsample_rate, start_idx, Nbins).  Refuse clearly rather than
    # fail with a TypeError in the middle of a calibration run (the reference only asks for torch >= 2.0).
    want = ["orig_hist", "orig_min", "orig_max", "update_hist", "update_min", "update_max"]
    have = list(inspect.signature(obs._combine_histograms).parameters)
    if have != want:
        raise NotImplementedError(
            f"HistogramObserver._combine_histograms{tuple(have)} of this torch release is not the one the device-histogram "
            f"replay was written against {tuple(want)}: calibrate this qconfig on the CPU (the reference's _calibrate), or "
            "use a min/max observer qconfig")
    if not (np.isfinite(float(x_min)) and np.isfinite(float(x_max))):
        raise RuntimeError(f"{where}: non-finite activation range ({float(x_min)}, {float(x_max)}) -- the reference's "
                           "HistogramObserver ignores infinities, the one-pass device histogram cannot")
    hist = torch.from_numpy(counts.astype(np.float32))
    cur_min, cur_max = obs.min_val, obs.max_val
    if cur_min == float("inf") or cur_max == float("-inf"):
        new_min, new_max = x_min, x_max                  # reset_histogram(x, x_min, x_max)
        combined = hist
    else:
        new_min, new_max = torch.min(cur_min, x_min), torch.max(cur_max, x_max)
        if new_min == cur_min and new_max == cur_max:
            combined = obs.histogram + hist
        else:
            combined = obs._combine_histograms(obs.histogram, cur_min, cur_max, hist, new_min, new_max)
    if (float(new_min), float(new_max)) != (float(used_range[0]), float(used_range[1])):
        raise RuntimeError(f"{where}: the device binned over {tuple(map(float, used_range))} but the observer's range at "
                           f"this call is {(float(new_min), float(new_max))} -- a later call of a shared observer widened "
                           "the range inside a batch, which the one-pass device histogram does not cover")
    obs.histogram.detach_().resize_(combined.shape)
    obs.histogram.copy_(combined)
    obs.min_val.detach_().resize_(new_min.shape)
    obs.min_val.copy_(new_min)
    obs.max_val.detach_().resize_(new_max.shape)
    obs.max_val.copy_(new_max)


def replay_observers(prepared, plan: List[Tuple[str, object]], stats: np.ndarray, num_tensors: int,
                     hists: Optional[np.ndarray] = None) -> None:
    """Feed the recorded per-batch statistics to the observers of ``prepared`` in the order the reference's calibration
    forward calls them (quantization/engines.py:130-133): batch by batch, node by node.  ``stats`` is the (min, max) log
    ``[records, points, 2]``; ``hists`` the histogram log ``[records, points, 2048]`` (needed by histogram observers:
    counts over the running range of the point's observer group after the batch, see ``point_groups``).

    A min/max observer updates its state from ``aminmax(x)`` alone, so it is simply called on the 2-element tensor
    ``[min, max]``; a histogram observer is advanced by ``_replay_histogram``."""
    from torch.ao.quantization.observer import HistogramObserver
    if stats.ndim != 3 or stats.shape[2] != 2:
        raise ValueError(f"expected [records, points, 2], got {stats.shape}")
    cols = [point_index(pt, num_tensors) for _, pt in plan]
    observers = [prepared.get_submodule(name) for name, _ in plan]
    if np.isnan(stats[:, cols]).any():
        raise RuntimeError("an observation point recorded no value (NaN activations or an unobserved tensor)")
    any_hist = any(isinstance(o, HistogramObserver) for o in observers)
    if any_hist:
        if hists is None or hists.shape[:2] != stats.shape[:2] or hists.shape[2] != HIST_BINS:
            raise ValueError("histogram observers need the histogram log [records, points, 2048]")
        groups = point_groups(prepared, plan, num_tensors)
        run = {}                                     # group -> running (min, max) after the current batch, as the device keeps it
    with torch.no_grad():
        for r, rec in enumerate(stats):
            if any_hist:
                for c in cols:
                    g = int(groups[c])
                    lo, hi = run.get(g, (np.float32(np.inf), np.float32(-np.inf)))
                    run[g] = (min(lo, rec[c, 0]), max(hi, rec[c, 1]))
            for (name, _), obs, c in zip(plan, observers, cols):
                if isinstance(obs, HistogramObserver):
                    _replay_histogram(obs, torch.tensor(rec[c, 0]), torch.tensor(rec[c, 1]), hists[r, c],
                                      run[int(groups[c])], f"record {r}, {name}")
                else:
                    obs(torch.tensor([rec[c, 0], rec[c, 1]], dtype=torch.float32))


class CalibrationEngine(B200HalfResNet):
    """FP16 engine over the prepared graph with one buffer per tensor and the device observer log switched on
    (``mode`` 1: per-batch (min, max); 2: also histograms over the running range of each point's observer group)."""

    def __init__(self, net: NetSpec, mode: int = 1, groups: Optional[np.ndarray] = None, **kw):
        super().__init__(net, **kw)
        self.mode = int(mode)
        self.set_option("keep_tensors", 1)
        self.set_option("observe", self.mode)
        self.num_points = int(self._lib.ievm_observer_points(self._handle))
        self.num_tensors = self.num_points - 2
        self.capacity = int(self._lib.ievm_observer_capacity(self._handle))
        self._groups = None
        if groups is not None:
            self.set_groups(groups)
        self._pending = 0
        self._stats: List[np.ndarray] = []
        self._hists: List[np.ndarray] = []

    def set_groups(self, groups: np.ndarray) -> None:
        g = np.ascontiguousarray(groups, dtype=np.int32)
        _lib.check(self._lib.ievm_observer_set_groups(self._handle, g.ctypes.data, int(g.size)), "ievm_observer_set_groups")
        self._groups = g

    def observe(self, x_f32: Optional[torch.Tensor] = None) -> None:
        """Append the statistics of the last forward to the device log (no synchronisation unless the log is full, in
        which case it is drained to the host first).  ``x_f32`` is the un-rounded f32 CUDA batch the forward's f16 input
        was cast from; without it the f16 input is observed."""
        if self._pending >= self.capacity:
            self._drain()
        ptr = None
        if x_f32 is not None:
            if not x_f32.is_cuda or x_f32.dtype != torch.float32 or not x_f32.is_contiguous():
                raise ValueError("x_f32 must be a contiguous float32 CUDA tensor")
            ptr = x_f32.data_ptr()
        stream = torch.cuda.current_stream(self.device_index).cuda_stream
        _lib.check(self._lib.ievm_observe(self._handle, ptr, stream), "ievm_observe")
        self._pending += 1

    def _drain(self) -> None:
        if self._pending == 0:
            return
        buf = np.empty((self._pending, self.num_points, 2), np.float32)
        n = self._lib.ievm_observer_read(self._handle, buf.ctypes.data, self._pending)
        if n != self._pending:
            _lib.check(n if n < 0 else -1, "ievm_observer_read")
        self._stats.append(buf)
        if self.mode == 2:
            hb = np.empty((self._pending, self.num_points, HIST_BINS), np.uint32)
            n = self._lib.ievm_observer_read_hist(self._handle, hb.ctypes.data, None, self._pending)
            if n != self._pending:
                _lib.check(n if n < 0 else -1, "ievm_observer_read_hist")
            self._hists.append(hb)
        _lib.check(self._lib.ievm_observer_clear(self._handle), "ievm_observer_clear")
        self._pending = 0

    def read_observations(self) -> np.ndarray:
        """[records, points, 2] float32 = (min, max) per ``observe`` call and observation point (synchronises)."""
        self._drain()
        return np.concatenate(self._stats) if self._stats else np.empty((0, self.num_points, 2), np.float32)

    def read_histograms(self) -> np.ndarray:
        """[records, points, 2048] uint32 (mode 2): ``torch.histc`` counts of every point over the running range of its
        observer group after that record's batch."""
        if self.mode != 2:
            raise RuntimeError("histograms are recorded with mode=2")
        self._drain()
        return np.concatenate(self._hists) if self._hists else np.empty((0, self.num_points, HIST_BINS), np.uint32)

    def reset_observations(self) -> None:
        self.set_option("observe", self.mode)
        if self._groups is not None:
            self.set_groups(self._groups)
        self._pending = 0
        self._stats, self._hists = [], []


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
    mode = observer_mode(prepared, plan)
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
                num_tensors = 1 + max(L.out_tensor for L in net.layers)
                eng = CalibrationEngine(net, mode=mode, groups=point_groups(prepared, plan, num_tensors), device=dev,
                                        max_batch=int(max_batch or images.shape[0]))
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
        replay_observers(prepared, plan, stats, eng.num_tensors, eng.read_histograms() if mode == 2 else None)
    finally:
        if eng is not None:
            eng.close()
    return stats if return_stats else None
