"""Six launches of the low-bit GEMM at the largest VAR-d30 mat_qkv shape (warm-up three, then three to capture with
ncu -k regex:gemm_codes -s 3 -c 3): groups of 128, row scales, row scales with CTA pairs (all 256-column tiles)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpqvar_b200 import _lib as L, lowbit          # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
m, n, k = 25600, 5760, 1920
x = torch.randn(m, k, device=dev, dtype=torch.float16)
w = torch.randn(n, k, device=dev) * 0.02
a, ww = lowbit.pack_codes(x, "e2m1"), lowbit.pack_codes(w, "e2m1")
ar, wr = lowbit.pack_codes(x, "e2m1", True), lowbit.pack_codes(w, "e2m1", True)
out = torch.empty(m, n, device=dev, dtype=torch.float16)
for _ in range(2):
    L.set_tunable("gemm_tile_n", 256)
    L.set_tunable("gemm_stages", 6)
    L.set_tunable("gemm_pair", 0)
    lowbit.linear_codes(a, ww, None, torch.float16, out)          # groups of 128, single CTAs
    lowbit.linear_codes(ar, wr, None, torch.float16, out)         # row scales, single CTAs
    L.set_tunable("gemm_pair", 1)
    lowbit.linear_codes(ar, wr, None, torch.float16, out)         # row scales, CTA pairs (cta_group::2)
    L.set_tunable("gemm_pair", -1)
    torch.cuda.synchronize()
print("ok")
