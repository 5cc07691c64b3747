#!/usr/bin/env python
"""Host-side cost per call of the Python layer (development aid): tiny tensors, many calls, wall clock."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpqvar_b200 import ops, quant_utils as Q
from fpqvar_b200.hotpath import seed42_sign_bits
x16 = torch.randn(100, 1920, device="cuda").half()
x32 = torch.randn(100, 1920, device="cuda")
sb = seed42_sign_bits()
s = torch.ones(1920, device="cuda")
def t(name, fn, n=3000):
    for _ in range(100): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); print(f"{name:50s} {(time.perf_counter()-t0)/n*1e6:7.2f} us/call")
t("Q.fp_quant_e2_per_group_cuda (fp16)", lambda: Q.fp_quant_e2_per_group_cuda(x16, 4, 128))
t("Q.fp_quant_e1m2_neg_e2m1_pos_per_group_cuda", lambda: Q.fp_quant_e1m2_neg_e2m1_pos_per_group_cuda(x16, 4, 128))
t("ops.transform_rotate_quant", lambda: ops.transform_rotate_quant(x32, s, sb, "e2m1"))
t("torch.empty_like only", lambda: torch.empty_like(x16))
t("x16 * 2 (one ATen kernel)", lambda: x16 * 2)
