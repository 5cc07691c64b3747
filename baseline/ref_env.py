"""MEASUREMENT INFRASTRUCTURE ONLY: imports the UNMODIFIED reference installed by baseline/install_ref.sh.

Nothing under fpqvar_b200/ imports this module; bench.py uses it for the reference legs (cpu_baseline,
--impl reference, reference_gpu_path, generation_reference_model) and falls back to the C port of the oracle
when no install is present.

The reference's files import two modules its tree does not ship: ``dist``
(``get_device()``, ``initialized()``; models*/var.py:9) and a top-level ``quant_utils``
(rotate_utils/rotation_utils.py:6, learnable_transformation/transform_model_utils.py:6; never used).  Both are
shimmed here, outside the reference tree, which stays byte-identical to /root/reference.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref", "FPQVAR")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models_fp_quant", "quant_utils.py"))


def why_unavailable() -> str:
    return f"{REF_ROOT} is not installed (run baseline/install_ref.sh where /root/reference exists)"


_loaded = {}


def load(device: str = "cpu"):
    """Returns a namespace with the reference's modules: qu (models_fp_quant_transform_rotate.quant_utils),
    qu0 (models_fp_quant.quant_utils), rotation_utils, transform_model_utils, build_vae_var (rotate variant),
    build_vae_var0 (models_fp_quant)."""
    if not available():
        raise RuntimeError(why_unavailable())
    if "ns" in _loaded:
        _loaded["dist"].get_device = lambda: device
        return _loaded["ns"]
    d = types.ModuleType("dist")
    d.get_device = lambda: device
    d.initialized = lambda: False
    d.get_rank = lambda: 0
    d.get_world_size = lambda: 1
    sys.modules.setdefault("dist", d)
    sys.modules.setdefault("quant_utils", types.ModuleType("quant_utils"))
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)              # packages + the reference's own quant_cuda extension
    ns = types.SimpleNamespace()
    ns.root = REF_ROOT
    ns.qu = importlib.import_module("models_fp_quant_transform_rotate.quant_utils")
    ns.qu0 = importlib.import_module("models_fp_quant.quant_utils")
    ns.rotation_utils = importlib.import_module("rotate_utils.rotation_utils")
    ns.transform_model_utils = importlib.import_module("learnable_transformation.transform_model_utils")
    ns.build_vae_var = importlib.import_module("models_fp_quant_transform_rotate").build_vae_var
    ns.build_vae_var0 = importlib.import_module("models_fp_quant").build_vae_var
    ns.quant_cuda = importlib.import_module("quant_cuda")
    _loaded["ns"] = ns
    _loaded["dist"] = sys.modules["dist"]
    return ns
