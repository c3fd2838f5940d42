"""Import alias for ``inference-efficient-vision-models_b200/`` (not a valid Python identifier)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "inference-efficient-vision-models_b200")
__path__.insert(0, _real)

from . import _lib  # noqa: E402
from .engine import B200HalfResNet, B200QuantizedResNet, PendingLogits, input_lut, kd_eval_loss, measure_mma_peak, pil_bilinear_coeffs  # noqa: E402
from .calibration import CalibrationEngine, calibrate, replay_observers
from .netdesc import NetSpec, from_converted, from_half_module, from_prepared, from_quantized_state_dict
from .pipeline import evaluate_accuracy, load_engine, measure_latency, save_engine  # noqa: E402

__all__ = ["B200QuantizedResNet", "B200HalfResNet", "PendingLogits", "measure_mma_peak", "kd_eval_loss", "input_lut", "pil_bilinear_coeffs", "NetSpec", "from_converted",
           "from_half_module", "from_quantized_state_dict", "from_prepared", "calibrate", "CalibrationEngine", "replay_observers", "load_engine", "save_engine", "evaluate_accuracy", "measure_latency", "_lib"]
