#!/usr/bin/env python
"""The reference's own GPU path on this B200, next to this library (evidence for profiles/, not a bench value).

Reference path = its UNMODIFIED CUDA extension (oracle/_ref/ref_quant_cuda*.so, built by oracle/build_ref.sh from
/root/reference/quant) driven by its Python glue restated operation for operation (tests/ref_glue.py:
models_fp_quant_transform_rotate/quant_utils.py:313-330, :415-452) plus, for the online rotation, the dense
`matmul(x.mul(s), Q)` under fp16 autocast of basic_var.py:263.  TEST INFRASTRUCTURE: uses oracle/_ref as the
comparator only."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_glue  # noqa: E402
from fpqvar_b200 import ops, quant_utils as Q, rotation_utils as R  # noqa: E402

ref = ref_glue.load_ref_ext()
if ref is None:
    raise SystemExit("oracle/_ref not built")
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = True          # evaluate_fp_quant_transform_rotate.py:172-175
torch.backends.cudnn.allow_tf32 = True


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3 / iters


rows, C = 25600, 1920
grid_e2m1 = Q.fp4_e2m1_grid.to(dev)
gneg = torch.tensor([-1.75, -1.5, -1.25, -1.0, -0.75, -0.5, -0.25, 0.0], device=dev)
gpos = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0], device=dev)
print(f"{'case':58s} {'reference path':>16s} {'this library':>16s} {'speed-up':>9s}")


def row(name, nbytes, f_ref, f_new):
    t_ref, t_new = timeit(f_ref), timeit(f_new)
    print(f"{name:58s} {nbytes / t_ref / 1e9:9.1f} GB/s {nbytes / t_new / 1e9:11.1f} GB/s {t_ref / t_new:8.1f}x")


x16 = torch.randn(rows, C, device=dev).half()
row("fp_quant_e2_per_group_cuda  fp16 [25600,1920] (proj in)", x16.numel() * 4,
    lambda: ref_glue.sym_group_cuda(ref.quant, x16, grid_e2m1), lambda: Q.fp_quant_e2_per_group_cuda(x16, 4, 128))
x32 = torch.randn(4096, 4096, device=dev)
row("fp_quant_e2_per_group_cuda  fp32 [4096,4096] (configs[0])", x32.numel() * 8,
    lambda: ref_glue.sym_group_cuda(ref.quant, x32, grid_e2m1), lambda: Q.fp_quant_e2_per_group_cuda(x32, 4, 128))
h16 = torch.nn.functional.gelu(torch.randn(rows, 4 * C, device=dev)).half()
row("fp_quant_e1m2_neg_e2m1_pos_per_group_cuda fp16 [25600,7680]", h16.numel() * 4,
    lambda: ref_glue.signsplit_group_cuda(ref.quant, h16, gneg, gpos), lambda: Q.fp_quant_e1m2_neg_e2m1_pos_per_group_cuda(h16, 4, 128))
# online site basic_var.py:263: dense rotation GEMM under autocast + .mul(s) + the quantizer
Qd = R.block_random_hadamard_matrix(C, 128, dev, 42).float()
s = torch.exp(torch.rand(C, device=dev) * 2 - 1)
xl = torch.randn(rows, C, device=dev)
bits = R.block_sign_bits()


def ref_online():
    with torch.autocast("cuda", dtype=torch.float16):
        x1 = torch.matmul(xl.mul(s), Qd)
    return ref_glue.sym_group_cuda(ref.quant, x1, grid_e2m1)


row("mul(s) + matmul(., Q) + fp_quant_e2 (mat_qkv in) [25600,1920]", xl.numel() * 6, ref_online,
    lambda: ops.transform_rotate_quant(xl, s, bits, "e2m1"))
sc, sh = torch.randn(100, 1, C, device=dev) * 0.3, torch.randn(100, 1, C, device=dev) * 0.5
x3 = xl.view(100, 256, C)


def ref_online_mod():
    with torch.autocast("cuda", dtype=torch.float16):
        x1 = torch.matmul(x3.mul(sc.add(1)).add_(sh).mul(s), Qd)
    return ref_glue.sym_group_cuda(ref.quant, x1, grid_e2m1)


row("adaLN modulate + mul(s) + matmul + fp_quant_e2 [100,256,1920]", xl.numel() * 6, ref_online_mod,
    lambda: ops.modulate_transform_rotate_quant(x3, sc, sh, s, bits, "e2m1"))
small = torch.randn(100, C, device=dev).half()
row("fp_quant_e2_per_group_cuda  fp16 [100,1920] (stage 0)", small.numel() * 4,
    lambda: ref_glue.sym_group_cuda(ref.quant, small, grid_e2m1), lambda: Q.fp_quant_e2_per_group_cuda(small, 4, 128))
