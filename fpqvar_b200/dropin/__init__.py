"""Modules importable under the reference's own names.

`quant_cuda` is a real extension module here (fpqvar_b200/dropin/quant_cuda*.so, built by
fpqvar_b200/csrc/build_torch_ext.sh from csrc/fpq_torch.cpp): the torch extension over the C ABI of libfpq_b200.so, with
the reference extension's entry point `quant(x, y) -> (z, idx)` (quant/quant.cpp:27-29).  Either

    sys.path.insert(0, ".../fpqvar_b200/dropin")      # ahead of the reference's quant/ build directory
    import quant_cuda                                  # the reference's quant_utils.py now runs on fpq_quant_grid

or

    import fpqvar_b200.dropin as dropin
    dropin.install()            # puts `quant_cuda`, `quant_utils`, `rotation_utils`, ... into sys.modules

Differences of `quant` a caller can observe, all deliberate: the launch goes to torch's CURRENT stream (the reference uses
the legacy default stream, quant_kernel.cu:52); `idx` -- which the reference allocates, zero-fills and never writes
(quant_kernel.cu:49,58) and which every one of its 91 call sites discards -- is a zero-stride expanded zero instead of a
fresh buffer; float64 input is rejected instead of being silently read as float32 (quant_kernel.cu:28)."""
import sys


def install(overwrite: bool = True) -> None:
    from . import quant_cuda
    from .. import block_rotation_utils, hadamard_utils, quant_utils, rotation_utils, transform_model_utils
    for name, mod in (("quant_cuda", quant_cuda), ("quant_utils", quant_utils), ("rotation_utils", rotation_utils),
                      ("block_rotation_utils", block_rotation_utils), ("hadamard_utils", hadamard_utils),
                      ("transform_model_utils", transform_model_utils)):
        if overwrite or name not in sys.modules:
            sys.modules[name] = mod
