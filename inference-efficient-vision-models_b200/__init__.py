"""B200-native (sm_100a) engine for the forward pass of the pruned ResNet-18 student (static INT8 or
FP16) and the ResNet-50 teacher of jaideepmurkute/Inference-Efficient-Vision-Models.

Import as ``ievm_b200`` (the directory name carries the reference's name and is not a valid Python
identifier; ``ievm_b200/__init__.py`` at the repo root aliases it).
"""
from . import _lib
from .engine import B200HalfResNet, B200QuantizedResNet, PendingLogits, input_lut, kd_eval_loss, measure_mma_peak, pil_bilinear_coeffs
from .calibration import CalibrationEngine, calibrate, replay_observers
from .netdesc import NetSpec, from_converted, from_half_module, from_prepared, from_quantized_state_dict
from .pipeline import evaluate_accuracy, load_engine, measure_latency, save_engine

__all__ = ["B200QuantizedResNet", "B200HalfResNet", "PendingLogits", "measure_mma_peak", "kd_eval_loss", "input_lut", "pil_bilinear_coeffs", "NetSpec", "from_converted",
           "from_half_module", "from_quantized_state_dict", "from_prepared", "calibrate", "CalibrationEngine", "replay_observers", "load_engine", "save_engine", "evaluate_accuracy", "measure_latency", "_lib"]
