"""The callers and data formats either side of the hot path.

* ``load_engine``: builds a B200 engine straight from the files the reference's stages write --
  ``model_static_int8.pth`` / ``model_fp16.pth`` (state dicts, quantization/main.py:306-308) and
  ``pruned_model.pth`` (whole pickled module, pruning/main.py:164-165, read back at quantization/main.py:124).
  Channel widths are recovered from the tensor shapes (structured pruning changes them per stage).
* ``evaluate_accuracy`` / ``measure_latency``: drop-ins for QuantizationEngine's helpers
  (quantization/engines.py:15-65) with the per-batch ``.item()`` synchronisation replaced by a device-side
  counter and host timers replaced by CUDA events.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .engine import B200HalfResNet, B200QuantizedResNet, _B200Engine
from .netdesc import DTYPE_F16, NetSpec


def _widths_from_float_state_dict(sd):
    """(block kind, stage widths, classes) of a torchvision ResNet state dict, pruned or not."""
    bottleneck = any(k.endswith("conv3.weight") for k in sd)
    widths = []
    for stage in (1, 2, 3, 4):
        last = "conv3" if bottleneck else "conv2"
        widths.append(int(sd[f"layer{stage}.0.{last}.weight"].shape[0]))
    return ("bottleneck" if bottleneck else "basic"), widths, int(sd["fc.weight"].shape[0])


def _module_from_float_state_dict(sd):
    """Rebuild the torchvision module a float / half state dict belongs to (widths from the tensor shapes)."""
    from .synthetic import make_student, make_teacher
    kind, widths, classes = _widths_from_float_state_dict(sd)
    if kind == "bottleneck":
        model = make_teacher(num_classes=classes)
    else:
        model = make_student(widths, num_classes=classes)
    model.load_state_dict({k: v.float() for k, v in sd.items()})
    return model.eval()


def load_engine(path_or_obj, max_batch: int = 256, device: Optional[int] = None) -> _B200Engine:
    """Engine from an on-disk artefact of the reference (or the already-loaded object):

    * quantized state dict (has ``conv1_input_scale_0``)        -> ``B200QuantizedResNet``
    * float / half ResNet state dict                            -> ``B200HalfResNet`` (FP16 cast, engines.py:84-93)
    * whole pickled ``nn.Module`` (``pruned_model.pth``)        -> ``B200HalfResNet`` of ``model.half()``
    * converted ``GraphModule``                                 -> ``B200QuantizedResNet``
    * ``.npz`` engine blob written by ``save_engine`` / ``NetSpec.save`` -> the engine it describes (no torch.ao needed)
    """
    obj = path_or_obj
    kw = dict(max_batch=max_batch, device=device)
    if isinstance(obj, (str, bytes)) or hasattr(obj, "__fspath__"):
        import os
        if os.fspath(obj).endswith(".npz" if isinstance(os.fspath(obj), str) else b".npz"):
            obj = NetSpec.load(obj)
        else:
            obj = torch.load(obj, map_location="cpu", weights_only=False)
    if isinstance(obj, NetSpec):
        return (B200HalfResNet if obj.dtype == DTYPE_F16 else B200QuantizedResNet)(obj, **kw)
    if isinstance(obj, dict):
        if "conv1_input_scale_0" in obj:
            return B200QuantizedResNet.from_quantized_state_dict(obj, **kw)
        if "conv1.weight" in obj and "bn1.weight" in obj:
            return B200HalfResNet.from_half_module(_module_from_float_state_dict(obj).half(), **kw)
        raise ValueError("unrecognised state dict: neither a converted static-INT8 graph nor a torchvision ResNet")
    if isinstance(obj, torch.fx.GraphModule) and hasattr(obj, "conv1_input_scale_0"):
        return B200QuantizedResNet.from_converted(obj, **kw)
    if isinstance(obj, torch.nn.Module):
        import copy
        return B200HalfResNet.from_half_module(copy.deepcopy(obj).eval().half(), **kw)
    raise TypeError(f"cannot build an engine from {type(obj).__name__}")


def save_engine(engine_or_net, path) -> None:
    """Export the flattened network of an engine (or a ``NetSpec``) as a single ``.npz`` blob for fast start:
    ``load_engine(path)`` rebuilds the same engine without the producer's torch.ao module."""
    net = engine_or_net if isinstance(engine_or_net, NetSpec) else engine_or_net.net
    net.save(path)


def evaluate_accuracy(model, data_loader, device="cuda") -> float:
    """quantization/engines.py:37-65 for a B200 engine: same protocol (FP16 engines get ``images.half()``,
    arg-max with the lowest index on ties), but correct/total accumulate in a device counter that is read
    once at the end instead of ``.item()`` after every batch.  Batches may be f32/f16 NCHW tensors or, for
    INT8 engines, uint8 NHWC images (``forward_u8``)."""
    lib = _lib.load()
    model.eval()
    try:
        param_dtype = next(model.parameters()).dtype
    except StopIteration:
        param_dtype = torch.float32
    dev = torch.device("cuda", model.device_index) if isinstance(model, _B200Engine) else torch.device(device)
    counters = torch.zeros(2, dtype=torch.int64, device=dev)
    host_correct = host_total = 0
    pending = []          # (PendingLogits, labels) of CPU batches in flight: at most two (ievm_submit_*_host)

    def settle(item):
        nonlocal host_correct, host_total
        logits, lab = item[0].result(), item[1]
        _, predicted = torch.max(logits.float(), 1)          # engines.py:61 (lowest index on ties)
        host_total += int(lab.shape[0])
        host_correct += int((predicted == lab).sum())

    with torch.no_grad():
        for images, labels in data_loader:
            if isinstance(model, _B200Engine) and not images.is_cuda and 0 < images.shape[0] <= model.max_batch:
                # the reference's loop hands over CPU batches (engines.py:52-60): pipelined host path -- the H2D copy of
                # this batch overlaps the forward of the previous one, and nothing synchronises per batch
                if images.dtype != torch.uint8 and param_dtype == torch.float16:
                    images = images.half()
                pending.append((model.submit(images), torch.as_tensor(labels).to(torch.int64)))
                if len(pending) == 2:
                    settle(pending.pop(0))
                continue
            images = images.to(dev, non_blocking=True)
            labels = torch.as_tensor(labels).to(dev, dtype=torch.int64, non_blocking=True)
            if images.dtype == torch.uint8:
                outputs = model.forward_u8(images)
            else:
                if param_dtype == torch.float16:
                    images = images.half()
                outputs = model(images)
            outputs = outputs.contiguous()
            dt = DTYPE_F16 if outputs.dtype == torch.float16 else 0
            if outputs.dtype not in (torch.float16, torch.float32):
                outputs = outputs.float()
            _lib.check(lib.ievm_count_correct(outputs.data_ptr(), dt, labels.data_ptr(), outputs.shape[0], outputs.shape[1],
                                              counters.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                       "ievm_count_correct")
        while pending:
            settle(pending.pop(0))
    correct, total = (int(v) for v in counters.cpu())
    return 100.0 * (correct + host_correct) / max(total + host_total, 1)


def measure_latency(model, input_dummy, num_runs: int = 100) -> float:
    """quantization/engines.py:15-35 (10 warm-up + ``num_runs`` timed calls, milliseconds per call), timed with
    CUDA events around the whole loop and a synchronise at the end -- the reference's host timer around
    asynchronous launches would measure enqueue time only."""
    model.eval()
    param_dtype = next(model.parameters()).dtype
    if input_dummy.dtype != param_dtype and input_dummy.dtype != torch.uint8:
        input_dummy = input_dummy.to(dtype=param_dtype)
    dev = torch.device("cuda", model.device_index)
    x = input_dummy.to(dev)
    call = model.forward_u8 if x.dtype == torch.uint8 else model
    with torch.no_grad():
        for _ in range(10):
            call(x)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        start.record()
        for _ in range(num_runs):
            call(x)
        end.record()
        torch.cuda.synchronize(dev)
    return start.elapsed_time(end) / num_runs
