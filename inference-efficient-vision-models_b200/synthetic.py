"""Synthetic stand-ins for the artefacts the reference's earlier stages would hand to the hot path.

The reference ships no checkpoints and no data (``*/output/.gitkeep``, ``data/.gitkeep``), so the
benchmark, the smoke test and the parity tests all start from seeded random models of the published
architecture.  Nothing here computes the hot path: these functions only *produce its inputs* with
stock torch / torchvision calls, exactly as the reference's stages do.

* ``make_student`` / ``make_teacher`` build the architectures the reference trains and then hands to
  its quantization stage: a torchvision ResNet-18 whose stage widths were shrunk by structured
  pruning (pruning/pruning_engine_structured.py:50-58 -- uniform local ratio, ``round_to=1``, ``fc``
  protected) with a 6-class head (quantization/q_config.py:17), and the ResNet-50 teacher with a
  6-class head (knowledge_distillation/utils.py:28-38).  BN statistics are randomised so that BN
  folding is non-trivial.
* ``static_quantize_fbgemm`` follows ``QuantizationEngine.static_quantize``
  (quantization/engines.py:95-121): default fbgemm qconfig mapping -> ``prepare_fx`` -> calibrate ->
  ``convert_fx``.  It additionally pins ``torch.backends.quantized.engine = "fbgemm"`` because the
  reference's stated backend is fbgemm and torch's default ("x86") is numerically different.
* ``cast_fp16`` follows ``QuantizationEngine.dynamic_quantize_fp16`` (quantization/engines.py:84-93).
"""
from __future__ import annotations

import copy
from typing import Sequence

import torch
import torch.nn as nn

PRUNED_WIDTHS = (57, 115, 230, 460)       # published run (9.02335 M params)
DEFAULT_CFG_WIDTHS = (60, 121, 243, 486)  # pruning/p_config.py:30 default ratio 0.05
UNPRUNED_WIDTHS = (64, 128, 256, 512)
NUM_CLASSES = 6                           # quantization/q_config.py:17


def _randomise(model: nn.Module, gen: torch.Generator) -> None:
    for mod in model.modules():
        if isinstance(mod, nn.Conv2d):
            fan_out = mod.out_channels * mod.kernel_size[0] * mod.kernel_size[1]
            std = (2.0 / fan_out) ** 0.5
            with torch.no_grad():
                mod.weight.copy_(torch.randn(mod.weight.shape, generator=gen) * std)
        elif isinstance(mod, nn.BatchNorm2d):
            c = mod.num_features
            with torch.no_grad():
                mod.running_mean.copy_(torch.randn(c, generator=gen) * 0.1)
                mod.running_var.copy_(torch.rand(c, generator=gen) + 0.5)
                mod.weight.copy_(torch.rand(c, generator=gen) + 0.5)
                mod.bias.copy_(torch.randn(c, generator=gen) * 0.1)
        elif isinstance(mod, nn.Linear):
            bound = 1.0 / mod.in_features ** 0.5
            with torch.no_grad():
                mod.weight.copy_((torch.rand(mod.weight.shape, generator=gen) * 2 - 1) * bound)
                mod.bias.copy_((torch.rand(mod.bias.shape, generator=gen) * 2 - 1) * bound)


def make_student(widths: Sequence[int] = PRUNED_WIDTHS, num_classes: int = NUM_CLASSES,
                 seed: int = 0) -> nn.Module:
    """torchvision ResNet-18 (BasicBlock x [2,2,2,2]) with per-stage widths ``widths``."""
    from torchvision.models.resnet import BasicBlock, ResNet

    model = ResNet(BasicBlock, [2, 2, 2, 2], num_classes=num_classes)
    w1, w2, w3, w4 = (int(w) for w in widths)
    model.inplanes = w1
    model.conv1 = nn.Conv2d(3, w1, kernel_size=7, stride=2, padding=3, bias=False)
    model.bn1 = nn.BatchNorm2d(w1)
    model.layer1 = model._make_layer(BasicBlock, w1, 2)
    model.layer2 = model._make_layer(BasicBlock, w2, 2, stride=2)
    model.layer3 = model._make_layer(BasicBlock, w3, 2, stride=2)
    model.layer4 = model._make_layer(BasicBlock, w4, 2, stride=2)
    model.fc = nn.Linear(w4, num_classes)
    _randomise(model, torch.Generator().manual_seed(seed))
    return model.eval()


def make_teacher(num_classes: int = NUM_CLASSES, seed: int = 0) -> nn.Module:
    """torchvision ResNet-50 with a ``num_classes`` head (knowledge_distillation/utils.py:28-38)."""
    from torchvision.models import resnet50

    model = resnet50(weights=None)
    model.fc = nn.Linear(model.fc.in_features, num_classes)
    _randomise(model, torch.Generator().manual_seed(seed))
    # Damp each residual branch's last BN so 16 stacked random blocks stay well inside fp16 range
    # (trained networks behave this way; torchvision offers zero_init_residual for the same reason).
    from torchvision.models.resnet import Bottleneck
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, Bottleneck):
                mod.bn3.weight.mul_(0.3)
    return model.eval()


def calibration_batches(n_batches: int = 2, batch: int = 8, seed: int = 1):
    gen = torch.Generator().manual_seed(seed)
    return [(torch.randn(batch, 3, 224, 224, generator=gen), torch.zeros(batch, dtype=torch.long))
            for _ in range(n_batches)]


def synthetic_images(n: int, seed: int = 7) -> torch.Tensor:
    return torch.randn(n, 3, 224, 224, generator=torch.Generator().manual_seed(seed))


def static_quantize_fbgemm(model: nn.Module, calib=None) -> nn.Module:
    """Restates quantization/engines.py:95-133 with the engine pinned to fbgemm."""
    from torch.ao.quantization import get_default_qconfig_mapping, quantize_fx

    torch.backends.quantized.engine = "fbgemm"
    calib = calibration_batches() if calib is None else calib
    work = copy.deepcopy(model).eval()           # prepare_fx mutates its input (main.py:179)
    qmap = get_default_qconfig_mapping("fbgemm")  # engines.py:103
    example_inputs = calib[0][0]                  # engines.py:105
    prepared = quantize_fx.prepare_fx(work, qmap, example_inputs)   # engines.py:109
    prepared.eval()
    with torch.no_grad():                         # engines.py:123-133
        for images, _ in calib:
            prepared(images.to("cpu"))
    return quantize_fx.convert_fx(prepared)       # engines.py:118


def minmax_qconfig_mapping():
    """The qconfig of the reference's stage-4 script (quantization/main.py:187-222): per-channel symmetric min/max weight
    observers, moving-average min/max activation observers over the full 0..255 range (no ``reduce_range``)."""
    from torch.ao.quantization import QConfig, QConfigMapping
    from torch.ao.quantization.observer import MovingAverageMinMaxObserver, PerChannelMinMaxObserver

    weight_observer = PerChannelMinMaxObserver.with_args(dtype=torch.qint8, qscheme=torch.per_channel_symmetric, ch_axis=0)
    activation_observer = MovingAverageMinMaxObserver.with_args(dtype=torch.quint8, qscheme=torch.per_tensor_affine,
                                                                averaging_constant=0.01)
    qc = QConfig(activation=activation_observer, weight=weight_observer)
    return (QConfigMapping().set_global(qc).set_object_type(nn.Conv2d, qc).set_object_type(nn.Linear, qc)
            .set_object_type(nn.ReLU, qc).set_object_type(nn.BatchNorm2d, qc))


def prepare_minmax(model: nn.Module, qconfig_mapping=None) -> nn.Module:
    """``prepare_fx`` as quantization/main.py:224-234 calls it (deep copy first: main.py:179).  The execution engine is
    pinned to fbgemm -- the reference's script selects qnnpack at :187, but the contract of this repo is fbgemm
    arithmetic and the observers do not depend on the engine."""
    from torch.ao.quantization import quantize_fx

    torch.backends.quantized.engine = "fbgemm"
    work = copy.deepcopy(model).eval()
    prepared = quantize_fx.prepare_fx(work, qconfig_mapping or minmax_qconfig_mapping(), (torch.randn(1, 3, 224, 224),))
    return prepared.eval()


def static_quantize_minmax(model: nn.Module, calib=None) -> nn.Module:
    """Restates the static_int8 branch of quantization/main.py:185-242 (CPU calibration loop at :236-239)."""
    from torch.ao.quantization import quantize_fx

    calib = calibration_batches() if calib is None else calib
    prepared = prepare_minmax(model)
    with torch.no_grad():
        for images, _ in calib:
            prepared(images.to("cpu"))
    return quantize_fx.convert_fx(prepared)


def cast_fp16(model: nn.Module) -> nn.Module:
    """Restates quantization/engines.py:84-93."""
    return copy.deepcopy(model).half().eval()
