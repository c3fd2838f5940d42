"""TEST INFRASTRUCTURE (oracle) -- generates tests/golden/*.npz from the reference's OWN code.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python oracle/gen_golden.py

It imports ``QuantizationEngine`` from /root/reference/quantization/engines.py and drives exactly
the calls the reference's stage-4 makes on the hot path:

  * ``static_quantize(model, calibration_loader, backend="fbgemm")`` (engines.py:95-121), then
    ``outputs = model(images)`` (engines.py:60) on CPU with the fbgemm engine;
  * ``dynamic_quantize_fp16(model)`` (engines.py:84-93), then ``model(images.half())``
    (engines.py:57-60).

For each synthetic model it stores the logits, every quantization parameter of the converted module
(so the GPU-box tests can prove their regenerated model is the same one), SHA-256 digests of every
graph node's uint8 activations, and a small raw slice of three activations.  The fixtures are the pin
for ``oracle/int8_forward.py`` / ``oracle/fp16_forward.py`` (tests/test_oracle.py).
"""
from __future__ import annotations

import hashlib
import logging
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/quantization")

from oracle import model_factory as mf  # noqa: E402

N_IMAGES = 8
N_IMAGES_FP16 = 4


def _digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def gen_int8(widths, tag):
    from engines import QuantizationEngine        # the reference's class, unmodified

    torch.backends.quantized.engine = "fbgemm"
    eng = QuantizationEngine(logging.getLogger("golden"))
    model = mf.make_student(widths)
    import copy
    gm = eng.static_quantize(copy.deepcopy(model), mf.calibration_batches(), backend="fbgemm")
    x = mf.synthetic_images(N_IMAGES)
    acts = {}
    for name, mod in gm.named_modules():
        if name and not list(mod.children()):
            mod.register_forward_hook(lambda m_, i, o, name=name: acts.__setitem__(name, o))
    with torch.no_grad():
        y = gm(x)
    out = {"logits": y.numpy(), "widths": np.array(widths), "n_images": np.array(N_IMAGES)}
    names, digests = [], []
    for name, o in acts.items():
        if getattr(o, "is_quantized", False):
            names.append(name)
            digests.append(_digest(o.int_repr().contiguous().numpy()))   # NCHW-contiguous bytes
            out[f"scale/{name}"] = np.float32(o.q_scale())
            out[f"zp/{name}"] = np.int32(o.q_zero_point())
    out["node_names"] = np.array(names)
    out["node_sha256"] = np.array(digests)
    out["in_scale"] = np.float32(float(gm.conv1_input_scale_0))
    out["in_zp"] = np.int32(int(gm.conv1_input_zero_point_0))
    for name in ("layer1.0.conv2", "layer3.0.downsample.0", "layer4.1.conv1"):
        out[f"slice/{name}"] = acts[name].int_repr()[0, :8, :4, :4].contiguous().numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"int8_{tag}.npz"), **out)
    print("wrote int8", tag, y[0])


def gen_int8_minmax(widths, tag):
    """The qconfig flavour of the reference's stage-4 SCRIPT (quantization/main.py:185-242: per-channel symmetric
    min/max weight observers, MovingAverageMinMaxObserver activations over 0..255).  That code is inline in main() and
    needs the dataset, so it cannot be imported; ``model_factory.static_quantize_minmax`` restates those lines (same
    torch.ao calls, same arguments) and every number below is produced by torch's own fbgemm operators on the module
    ``convert_fx`` returns.  This is the flavour on which quantized::add_relu's fused dequantisation matters."""
    torch.backends.quantized.engine = "fbgemm"
    gm = mf.static_quantize_minmax(mf.make_student(widths))
    x = mf.synthetic_images(N_IMAGES)
    acts = {}
    for name, mod in gm.named_modules():
        if name and not list(mod.children()):
            mod.register_forward_hook(lambda m_, i, o, name=name: acts.__setitem__(name, o))
    with torch.no_grad():
        y = gm(x)
    out = {"logits": y.numpy(), "widths": np.array(widths), "n_images": np.array(N_IMAGES),
           "in_scale": np.float32(float(gm.conv1_input_scale_0)), "in_zp": np.int32(int(gm.conv1_input_zero_point_0))}
    names, digests = [], []
    for name, o in acts.items():
        if getattr(o, "is_quantized", False):
            names.append(name)
            digests.append(_digest(o.int_repr().contiguous().numpy()))
    out["node_names"] = np.array(names)
    out["node_sha256"] = np.array(digests)
    for i in range(8):                 # quantized.add_relu outputs are graph functions, not modules: read their qparams
        sfx = "" if i == 0 else f"_{i}"
        li, bi = 1 + i // 2, i % 2
        out[f"scale/add_relu{sfx}"] = np.float32(float(getattr(gm, f"layer{li}_{bi}_relu_scale_0")))
        out[f"zp/add_relu{sfx}"] = np.int32(int(getattr(gm, f"layer{li}_{bi}_relu_zero_point_0")))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"int8_minmax_{tag}.npz"), **out)
    print("wrote int8 minmax", tag, y[0])


def gen_fp16(widths, tag):
    from engines import QuantizationEngine

    eng = QuantizationEngine(logging.getLogger("golden"))
    model = mf.make_student(widths)
    m16 = eng.dynamic_quantize_fp16(model)
    x = mf.synthetic_images(N_IMAGES_FP16)
    with torch.no_grad():
        y16 = m16(x.half())
        y32 = model(x)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"fp16_{tag}.npz"),
                        logits_fp16=y16.float().numpy(), logits_fp32=y32.numpy(),
                        widths=np.array(widths), n_images=np.array(N_IMAGES_FP16))
    print("wrote fp16", tag, y16[0])


def gen_teacher():
    from engines import QuantizationEngine

    eng = QuantizationEngine(logging.getLogger("golden"))
    model = mf.make_teacher()
    m16 = eng.dynamic_quantize_fp16(model)
    x = mf.synthetic_images(2)
    with torch.no_grad():
        y16 = m16(x.half())
        y32 = model(x)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fp16_teacher_r50.npz"),
                        logits_fp16=y16.float().numpy(), logits_fp32=y32.numpy(), n_images=np.array(2))
    print("wrote teacher", y16[0])


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    gen_int8(mf.PRUNED_WIDTHS, "w57")
    gen_int8(mf.DEFAULT_CFG_WIDTHS, "w60")
    gen_int8(mf.UNPRUNED_WIDTHS, "w64")
    gen_int8_minmax(mf.PRUNED_WIDTHS, "w57")
    gen_fp16(mf.PRUNED_WIDTHS, "w57")
    gen_teacher()
