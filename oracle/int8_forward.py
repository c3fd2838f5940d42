"""TEST INFRASTRUCTURE (oracle) -- closed-form CPU restatement of the INT8 hot path.

Not product code: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may
import it.  The product path (the CUDA engine) never calls into this file.

The reference (jaideepmurkute/Inference-Efficient-Vision-Models) holds *no* arithmetic of its own for
this path: ``outputs = model(images)`` (quantization/engines.py:27,31,60; quantization/main.py:287)
runs the FX-converted ``GraphModule`` created at quantization/engines.py:118, whose nodes dispatch to
PyTorch's quantized CPU operators -> FBGEMM.  That dependency is un-vendored and unpinned
(requirements.txt:1 ``torch>=2.0.0``; this image: torch 2.11.0+cu128, FBGEMM bundled in the wheel).
This module restates the *published semantics* of those operators in plain integer / float32
arithmetic so that int32 accumulators (which torch never exposes) are available to the parity tests:

  quantize_per_tensor   q = clamp(rne(x * (1/s)) + zp, 0, 255)              (ATen quantize_val / fbgemm Quantize)
  quantized::conv2d[_relu], quantized::linear
                        acc = sum (xq - x_zp) * wq          (int32; zero padding == x_zp)
                        v   = (float(acc) + bias/(x_s*w_s[c])) * ((x_s*w_s[c]) / out_s)   (fbgemm ReQuantizeOutput, float bias)
                        q   = clamp(rne(v) + out_zp, lo, 255), lo = out_zp if relu else 0
  quantized max_pool2d  integer max over the window (padding ignored), qparams pass through
  quantized::add_relu   a = fma(a_s, aq, fl(a_s * -a_zp)) ; b = fma(b_s, bq, fl(b_s * -b_zp))   (ATen's vector path)
                        q = clamp(rne(max(a + b, 0) * (1/s)) + zp, 0, 255)
  adaptive_avg_pool2d   q = clamp(rne(float(sum) / count), 0, 255)           (zp == 0 on this path)
  dequantize            y = (q - zp) * s

PARITY PIN: the reference ships no tests, fixtures or golden vectors ("parity unpinned" by the
reference itself).  The pin used instead is the reference's own producer
(``QuantizationEngine.static_quantize``) imported from /root/reference and run on CPU with the fbgemm
engine: ``oracle/gen_golden.py`` records its logits and per-node activation digests in
``tests/golden/`` and ``tests/test_oracle.py`` checks this restatement against them bit-for-bit and
against the live torch fbgemm operators node by node.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class QConv:
    name: str
    w: np.ndarray            # int8 [Cout, Cin, kh, kw]
    w_scale: np.ndarray      # float32 [Cout]
    bias: np.ndarray         # float32 [Cout]
    stride: int
    pad: int
    relu: bool
    out_scale: float
    out_zp: int


@dataclass
class QBlock:
    conv1: QConv
    conv2: QConv
    down: Optional[QConv]
    add_scale: float
    add_zp: int


@dataclass
class QNet:
    in_scale: float
    in_zp: int
    stem: QConv
    blocks: List[QBlock]
    fc_w: np.ndarray         # int8 [classes, Cin]
    fc_w_scale: np.ndarray   # float32 [classes]
    fc_bias: np.ndarray      # float32 [classes]
    fc_scale: float
    fc_zp: int
    trace: Dict[str, np.ndarray] = field(default_factory=dict)


def _qconv_from_module(name: str, mod) -> QConv:
    w = mod.weight()
    assert w.qscheme() in (torch.per_channel_affine, torch.per_channel_symmetric)
    assert int(w.q_per_channel_zero_points().abs().max()) == 0
    b = mod.bias()
    cout = w.shape[0]
    return QConv(
        name=name,
        w=w.int_repr().numpy().copy(),
        w_scale=w.q_per_channel_scales().to(torch.float32).numpy().copy(),
        bias=(b.detach().float().numpy().copy() if b is not None else np.zeros(cout, np.float32)),
        stride=int(mod.stride[0]), pad=int(mod.padding[0]),
        relu=type(mod).__name__ == "ConvReLU2d",
        out_scale=float(mod.scale), out_zp=int(mod.zero_point))


def extract_qnet(gm) -> QNet:
    """Read a converted ResNet-18-topology GraphModule (quantization/engines.py:118 output)."""
    mods = dict(gm.named_modules())
    blocks = []
    for li in range(1, 5):
        for bi in range(2):
            p = f"layer{li}.{bi}"
            down = mods.get(f"{p}.downsample.0")
            blocks.append(QBlock(
                conv1=_qconv_from_module(f"{p}.conv1", mods[f"{p}.conv1"]),
                conv2=_qconv_from_module(f"{p}.conv2", mods[f"{p}.conv2"]),
                down=_qconv_from_module(f"{p}.downsample.0", down) if down is not None else None,
                add_scale=float(getattr(gm, f"layer{li}_{bi}_relu_scale_0")),
                add_zp=int(getattr(gm, f"layer{li}_{bi}_relu_zero_point_0"))))
    fcw, fcb = mods["fc"]._packed_params._weight_bias()
    return QNet(
        in_scale=float(gm.conv1_input_scale_0), in_zp=int(gm.conv1_input_zero_point_0),
        stem=_qconv_from_module("conv1", mods["conv1"]), blocks=blocks,
        fc_w=fcw.int_repr().numpy().copy(),
        fc_w_scale=fcw.q_per_channel_scales().to(torch.float32).numpy().copy(),
        fc_bias=fcb.detach().float().numpy().copy(),
        fc_scale=float(mods["fc"].scale), fc_zp=int(mods["fc"].zero_point))


# ----------------------------------------------------------------------------- arithmetic

def _rne(x: np.ndarray) -> np.ndarray:
    return np.rint(x)          # IEEE round-half-to-even, as cvtps2dq / nearbyint


def quantize_input(x: np.ndarray, scale: float, zp: int) -> np.ndarray:
    inv = np.float32(1.0) / np.float32(scale)
    q = _rne(x.astype(np.float32) * inv) + np.float32(zp)
    return np.clip(q, 0, 255).astype(np.uint8)


def conv_acc(xq: np.ndarray, x_zp: int, w: np.ndarray, stride: int, pad: int) -> np.ndarray:
    """Exact integer convolution of (xq - x_zp) with w; NCHW u8 in, NCHW int32 out."""
    xs = torch.from_numpy(xq.astype(np.float64) - float(x_zp))
    ws = torch.from_numpy(w.astype(np.float64))
    acc = F.conv2d(xs, ws, None, stride=stride, padding=pad)   # exact: |acc| << 2^53
    return acc.numpy().astype(np.int64).astype(np.int32)


def requant(acc: np.ndarray, x_scale: float, w_scale: np.ndarray, bias: np.ndarray,
            out_scale: float, out_zp: int, relu: bool, ch_axis: int = 1) -> np.ndarray:
    shp = [1] * acc.ndim
    shp[ch_axis] = -1
    atw = (np.float32(x_scale) * w_scale.astype(np.float32)).astype(np.float32)
    bdiv = (bias.astype(np.float32) / atw).astype(np.float32).reshape(shp)
    mult = (atw / np.float32(out_scale)).astype(np.float32).reshape(shp)
    v = (acc.astype(np.float32) + bdiv).astype(np.float32) * mult
    q = _rne(v.astype(np.float32)) + np.float32(out_zp)
    lo = out_zp if relu else 0
    return np.clip(q, lo, 255).astype(np.uint8)


def maxpool3x3s2(xq: np.ndarray) -> np.ndarray:
    t = torch.from_numpy(xq.astype(np.int32)).float()
    return F.max_pool2d(t, 3, 2, 1).numpy().astype(np.uint8)    # -inf padding == padding ignored


def _dequant_fma(q: np.ndarray, scale: float, zp: int) -> np.ndarray:
    """ATen's vectorised dequantisation inside qadd (aten/src/ATen/native/quantized/cpu/kernels/QuantizedOpKernels.cpp
    ``qadd_kernel``: ``Vectorized<c10::quint8>::dequantize(scale, zero_point, scale_zp_premul)``):
    ``fma(scale, float(q), fl32(scale * -zp))`` -- one rounding on top of a pre-rounded product, not ``(q - zp) * scale``.
    The two forms differ by an ulp on many elements and by one output LSB on about 1e-5 of them when both operands have a
    large zero point (observed: layer4.0's add of the quantization/main.py qconfig flavour; tests/test_calibration.py).
    float64 holds ``q * scale + premul`` exactly (8 + 24 bits against 24 bits), so one cast to float32 is the fma."""
    s32 = np.float32(scale)
    premul = np.float32(s32 * np.float32(-zp))
    return (q.astype(np.float64) * np.float64(s32) + np.float64(premul)).astype(np.float32)


def add_relu(aq, a_scale, a_zp, bq, b_scale, b_zp, out_scale, out_zp) -> np.ndarray:
    a = _dequant_fma(aq, a_scale, a_zp)
    b = _dequant_fma(bq, b_scale, b_zp)
    s = np.maximum((a + b).astype(np.float32), np.float32(0))
    inv = np.float32(1.0) / np.float32(out_scale)
    q = _rne((s * inv).astype(np.float32)) + np.float32(out_zp)
    return np.clip(q, 0, 255).astype(np.uint8)


def avgpool(xq: np.ndarray) -> np.ndarray:
    n, c, h, w = xq.shape
    s = xq.astype(np.int32).sum(axis=(2, 3)).astype(np.float32)
    q = _rne((s / np.float32(h * w)).astype(np.float32))
    return np.clip(q, 0, 255).astype(np.uint8)


def forward(net: QNet, x: np.ndarray, keep: bool = False) -> np.ndarray:
    """x: float32 NCHW.  Returns float32 logits [N, classes].  With keep=True, ``net.trace`` holds
    every node's u8 activations (NCHW) and int32 accumulators keyed by the reference graph's names."""
    tr: Dict[str, np.ndarray] = {}

    def run_conv(c: QConv, xq, x_scale, x_zp):
        acc = conv_acc(xq, x_zp, c.w, c.stride, c.pad)
        out = requant(acc, x_scale, c.w_scale, c.bias, c.out_scale, c.out_zp, c.relu)
        if keep:
            tr[c.name + ":acc"] = acc
            tr[c.name] = out
        return out

    xq = quantize_input(x, net.in_scale, net.in_zp)
    if keep:
        tr["quantize_per_tensor"] = xq
    cur = run_conv(net.stem, xq, net.in_scale, net.in_zp)
    cur_s, cur_zp = net.stem.out_scale, net.stem.out_zp
    cur = maxpool3x3s2(cur)
    if keep:
        tr["maxpool"] = cur
    for bi, blk in enumerate(net.blocks):
        t = run_conv(blk.conv1, cur, cur_s, cur_zp)
        a = run_conv(blk.conv2, t, blk.conv1.out_scale, blk.conv1.out_zp)
        if blk.down is not None:
            r = run_conv(blk.down, cur, cur_s, cur_zp)
            r_s, r_zp = blk.down.out_scale, blk.down.out_zp
        else:
            r, r_s, r_zp = cur, cur_s, cur_zp
        cur = add_relu(a, blk.conv2.out_scale, blk.conv2.out_zp, r, r_s, r_zp, blk.add_scale, blk.add_zp)
        cur_s, cur_zp = blk.add_scale, blk.add_zp
        if keep:
            tr["add_relu" + ("" if bi == 0 else f"_{bi}")] = cur
    pooled = avgpool(cur)
    if keep:
        tr["avgpool"] = pooled
    acc = (pooled.astype(np.int64) - cur_zp) @ net.fc_w.astype(np.int64).T
    acc = acc.astype(np.int32)
    q = requant(acc, cur_s, net.fc_w_scale, net.fc_bias, net.fc_scale, net.fc_zp, False, ch_axis=1)
    if keep:
        tr["fc:acc"] = acc
        tr["fc"] = q
        net.trace = tr
    return ((q.astype(np.float32) - np.float32(net.fc_zp)) * np.float32(net.fc_scale)).astype(np.float32)
