"""Modules importable under the reference's own names.

    import fpqvar_b200.dropin as dropin
    dropin.install()            # puts `quant_cuda`, `quant_utils`, ... into sys.modules
    import quant_cuda           # -> fpqvar_b200.dropin.quant_cuda
    quant_cuda.quant(x, grid)

or put this directory on sys.path ahead of the reference's quant/ build directory."""
import sys


def install(overwrite: bool = True) -> None:
    from . import quant_cuda
    from .. import block_rotation_utils, hadamard_utils, quant_utils, rotation_utils, transform_model_utils
    for name, mod in (("quant_cuda", quant_cuda), ("quant_utils", quant_utils), ("rotation_utils", rotation_utils),
                      ("block_rotation_utils", block_rotation_utils), ("hadamard_utils", hadamard_utils),
                      ("transform_model_utils", transform_model_utils)):
        if overwrite or name not in sys.modules:
            sys.modules[name] = mod
