"""TEST INFRASTRUCTURE (oracle) -- re-exports the synthetic model / input producers.

The producers live in the package (``ievm_b200.synthetic``) because the benchmark's product arm needs
them too and may not import ``oracle/``; the oracle and the golden-vector generator use the very same
functions so that every tier sees the same seeded models.
"""
from ievm_b200.synthetic import *  # noqa: F401,F403
from ievm_b200.synthetic import (DEFAULT_CFG_WIDTHS, NUM_CLASSES, PRUNED_WIDTHS, UNPRUNED_WIDTHS,  # noqa: F401
                                 calibration_batches, cast_fp16, make_student, make_teacher, minmax_qconfig_mapping,
                                 prepare_minmax, static_quantize_minmax,
                                 static_quantize_fbgemm, synthetic_images)
