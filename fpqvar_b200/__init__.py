"""fpqvar_b200 -- B200-native (sm_100a) floating-point fake-quantization hot path of FPQVAR.

Layout:
  csrc/               hand-written CUDA kernels + the C ABI (include/fpq_b200.h) -> libfpq_b200.so
  ops.py              tensor-level operators over the C ABI
  quant_utils.py      mirror of the reference's models_fp_quant*/quant_utils.py (functions + classes)
  rotation_utils.py   mirror of rotate_utils/rotation_utils.py (block Hadamard) on the fused kernels
  transform_model_utils.py  mirror of learnable_transformation/transform_model_utils.py
  search.py           batched format scoring + search loop (search/search_fp*_format.py)
  dropin/             modules importable under the reference's own names (quant_cuda, quant_utils, ...)
  var_workload.py     shapes of the hot path inside one VAR generation pass (what bench.py replays)
  hotpath.py          graph-replayable device path and the host-buffer pipeline
"""
from . import _lib  # noqa: F401
from . import ops  # noqa: F401

__all__ = ["ops"]
