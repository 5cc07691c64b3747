"""Generate tests/golden/reference_mixed_plans.json: which format every quantized linear gets under the reference's
per-layer ("mixed datatype") variants of quantize_VAR.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_mixed.py

The reference's own functions run on a 30-block toy model built from its own `AdaLNSelfAttn` class (width 128), with
`QuantizedLinear.from_float` / `QuantizedLinear_fc2.from_float` replaced by a recorder, so the fixture holds the exact
sequence of (module path, class, keyword arguments) the reference issues:
  models_fp_quant/quant_utils.py:1256-1341  quantize_VAR_mixed_fp4_datatype
  models_fp_quant/quant_utils.py:1344-1432  quantize_VAR_mixed_fp6_datatype
  models_fp_quant_rotate/quant_utils.py:982-1067  quantize_VAR_use_different_datatype
"""
import importlib
import json
import os
import sys
from functools import partial

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402

KW = dict(weight_quant="per_group", act_quant="per_group", w_bit=4, a_bit=4, act_quant_sym=True, fc2_act_log2_quant=False,
          activation_fp_quant=True, weight_fp_quant=True, act_fp_type="fp_e2", weight_fp_type="fp_e2",
          fc2_fp_type="fp_e1m2_neg_e2m1_pos")


def run(pkg, fn_name, kw):
    qu = importlib.import_module(f"{pkg}.quant_utils")
    bv = importlib.import_module(f"{pkg}.basic_var")

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.blocks = torch.nn.ModuleList(
                bv.AdaLNSelfAttn(block_idx=i, last_drop_p=0, embed_dim=128, cond_dim=128, shared_aln=False,
                                 norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_heads=2, mlp_ratio=1.0,
                                 attn_l2_norm=True, flash_if_available=False, fused_if_available=False) for i in range(30))

    model = Toy()
    paths = {id(m): n for n, m in model.named_modules()}
    calls = []

    def recorder(cls_name):
        def from_float(module, **kwargs):
            calls.append({"module": paths[id(module)], "class": cls_name,
                          "kwargs": {k: kwargs[k] for k in sorted(kwargs)}})
            return module
        return staticmethod(from_float)

    qu.QuantizedLinear.from_float = recorder("QuantizedLinear")
    qu.QuantizedLinear_fc2.from_float = recorder("QuantizedLinear_fc2")
    getattr(qu, fn_name)(model, **kw)
    return calls


def main():
    MG._install_shims()
    out = {"kwargs": KW,
           "quantize_VAR_mixed_fp4_datatype": run("models_fp_quant", "quantize_VAR_mixed_fp4_datatype", KW),
           "quantize_VAR_mixed_fp6_datatype": run("models_fp_quant", "quantize_VAR_mixed_fp6_datatype",
                                                  dict(KW, w_bit=6, a_bit=6, act_fp_type="fp6_e2m3", weight_fp_type="fp6_e2m3",
                                                       fc2_fp_type="fp6_int_neg_e2m3_pos")),
           "quantize_VAR_use_different_datatype": run("models_fp_quant_rotate", "quantize_VAR_use_different_datatype", KW),
           # models_fp_quant_rotate/quant_utils.py:894-979: this package's quantize_VAR also quantizes every block's ada_lin[1]
           "models_fp_quant_rotate.quantize_VAR": run("models_fp_quant_rotate", "quantize_VAR", KW),
           # the README package: ada_lin stays FP (the branch is commented out, qu.py:1147-1155)
           "models_fp_quant_transform_rotate.quantize_VAR": run("models_fp_quant_transform_rotate", "quantize_VAR", KW)}
    path = os.path.join(HERE, "reference_mixed_plans.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print(f"wrote {path}: " + ", ".join(f"{k}: {len(v)} calls" for k, v in out.items() if k != "kwargs"))


if __name__ == "__main__":
    main()
