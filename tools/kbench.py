#!/usr/bin/env python
"""Kernel micro-benchmark (development aid): times the hot kernels alone at the largest
VAR-d30 stage shapes with rotating buffers (so that L2 cannot serve repeats), CUDA events,
prints algorithmic GB/s.  FPQ_LIB_PATH selects an alternative build of libfpq_b200.so."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fpqvar_b200 import _lib as L  # noqa: E402

if os.environ.get("FPQ_LIB_PATH"):
    L.LIB_PATH = os.environ["FPQ_LIB_PATH"]
from fpqvar_b200 import ops  # noqa: E402
from fpqvar_b200.hotpath import seed42_sign_bits  # noqa: E402

lib = L.lib()
for kv in os.environ.get("FPQ_TUNABLES", "").split(","):        # e.g. FPQ_TUNABLES=smem_kb=100,pdl=0
    if kv:
        k, v = kv.split("=")
        L.set_tunable(k, int(v))
dev = torch.device("cuda")
st = torch.cuda.current_stream().cuda_stream
NBUF = int(os.environ.get("KB_NBUF", "6"))
ITERS = int(os.environ.get("KB_ITERS", "30"))
ONLY = os.environ.get("KB_ONLY")


def smi():
    import subprocess
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader"],
                              capture_output=True, text=True).stdout.strip()
    except Exception:
        return ""


def timeit(fn, nbytes):
    for i in range(NBUF):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(ITERS):
        fn(i % NBUF)
    e1.record()
    if ITERS >= 1000:
        print("   under load:", smi())
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 1e3 / ITERS
    return nbytes / t / 1e9, t * 1e6


def main():
    rows = int(os.environ.get("ROWS", "25600"))
    C = 1920
    flag = torch.zeros(2, dtype=torch.int32, device=dev)      # {flag, ticket}
    sb = seed42_sign_bits()
    smooth = torch.exp(torch.rand(C, device=dev) * 2 - 1)
    class R(dict):
        def __setitem__(self, k, v):
            dict.__setitem__(self, k, v)
    res = R()
    def want(name):
        return ONLY is None or any(o in name for o in ONLY.split("|"))
    # sign-split fc2 input
    x = [torch.nn.functional.gelu(torch.randn(rows, 4 * C, device=dev)).half() for _ in range(NBUF)]
    o = [torch.empty_like(t) for t in x]
    n = rows * 4 * C
    if want("signsplit f16 (fc2)"):
        res["signsplit f16 (fc2)"] = timeit(lambda i: lib.fpq_fake_quant_signsplit(x[i].data_ptr(), o[i].data_ptr(), n // 128, 128, 1, 1, 0, 0, 0, None, st), n * 4)
    if want("signsplit f16 +clip"):
        res["signsplit f16 +clip"] = timeit(lambda i: lib.fpq_fake_quant_signsplit(x[i].data_ptr(), o[i].data_ptr(), n // 128, 128, 1, 1, 0, 0, 2, flag.data_ptr(), st), n * 4)
    if want("gelu+signsplit f16 fused"):
        xg = [torch.randn(rows, 4 * C, device=dev).half() for _ in range(NBUF)]
        res["gelu+signsplit f16 fused (8 B/elem counted)"] = timeit(lambda i: lib.fpq_gelu_fake_quant_signsplit(xg[i].data_ptr(), o[i].data_ptr(), n // 128, 0, 2, flag.data_ptr(), st), n * 8)
        tmp = [torch.empty_like(t) for t in xg]
        def unfused(i):
            tmp[i] = torch.nn.functional.gelu(xg[i], approximate="tanh")
            lib.fpq_fake_quant_signsplit(tmp[i].data_ptr(), o[i].data_ptr(), n // 128, 128, 1, 1, 0, 0, 2, flag.data_ptr(), st)
        res["gelu (ATen) then signsplit, 2 launches (8 B/elem)"] = timeit(unfused, n * 8)
        res["gelu (ATen) alone (4 B/elem)"] = timeit(lambda i: torch.nn.functional.gelu(xg[i], approximate="tanh"), n * 4)
        del xg, tmp
    if want("group e2m1 f16 (4C)"):
        res["group e2m1 f16 (4C)"] = timeit(lambda i: lib.fpq_fake_quant(x[i].data_ptr(), o[i].data_ptr(), n // 128, 128, 1, 1, 0, 0, 0, st), n * 4)
    if want("group e2m3 f16 (4C)"):
        res["group e2m3 f16 (4C)"] = timeit(lambda i: lib.fpq_fake_quant(x[i].data_ptr(), o[i].data_ptr(), n // 128, 128, 1, 1, 3, 0, 0, st), n * 4)
    # per-token / per-channel rows (FP6 README configs: rows of 9216 = 4C of VAR-d36)
    nr = (rows * 4 * C) // 9216
    xr = [t.view(-1)[: nr * 9216] for t in x]
    orr = [t.view(-1)[: nr * 9216] for t in o]
    if want("rows e2m3 f16 (per_token 9216)"):
        res["rows e2m3 f16 (per_token 9216)"] = timeit(lambda i: lib.fpq_fake_quant(xr[i].data_ptr(), orr[i].data_ptr(), nr, 9216, 1, 1, 3, 0, 0, st), nr * 9216 * 4)
    if want("rows int_neg_e2m3_pos f16 (per_token 9216)"):
        res["rows int_neg_e2m3_pos f16 (per_token 9216)"] = timeit(lambda i: lib.fpq_fake_quant_signsplit(xr[i].data_ptr(), orr[i].data_ptr(), nr, 9216, 1, 1, 1, 0, 0, None, st), nr * 9216 * 4)
    xk = [t.view(-1, 64) for t in x]
    if want("rows e2m3 f16 (KV rows of 64)"):
        res["rows e2m3 f16 (KV rows of 64)"] = timeit(lambda i: lib.fpq_fake_quant(x[i].data_ptr(), o[i].data_ptr(), n // 64, 64, 1, 1, 3, 0, 0, st), n * 4)
    if want("score 3 fmts f16 (read-only)"):
        res["score 3 fmts f16 (read-only)"] = timeit(lambda i: ops.score_formats(x[i], ["e2m1", "e1m2", "e3m0"]), n * 2)
    if want("score 4 fmts f16 incl split"):
        res["score 4 fmts f16 incl split"] = timeit(lambda i: ops.score_formats(x[i], ["e2m1", "e1m2", "e3m0", "e1m2_neg_e2m1_pos"]), n * 2)
    del x, o, xr, orr, xk
    # rotate + quant (mat_qkv / fc1 input), 4 row-blocks to get a comparable byte volume
    r4 = rows * 4
    xf = [torch.randn(r4, C, device=dev) for _ in range(NBUF)]
    of = [torch.empty(r4, C, device=dev, dtype=torch.float16) for _ in range(NBUF)]
    n = r4 * C
    if want("rotate+quant f32->f16"):
        res["rotate+quant f32->f16"] = timeit(lambda i: lib.fpq_transform_rotate_quant(xf[i].data_ptr(), smooth.data_ptr(), sb, of[i].data_ptr(), None, r4, C, 0, st), n * 6)
    if want("rotate only f32->f16"):
        res["rotate only f32->f16"] = timeit(lambda i: lib.fpq_transform_rotate_quant(xf[i].data_ptr(), smooth.data_ptr(), sb, of[i].data_ptr(), None, r4, C, -1, st), n * 6)
    if want("adaLN+rotate+quant f32->f16"):
        B = 100
        sc = torch.randn(B, C, device=dev) * 0.3
        sh = torch.randn(B, C, device=dev) * 0.5
        res["adaLN+rotate+quant f32->f16"] = timeit(lambda i: lib.fpq_modulate_transform_rotate_quant(
            xf[i].data_ptr(), sc.data_ptr(), sh.data_ptr(), r4 // B, smooth.data_ptr(), sb, of[i].data_ptr(), None, r4, C, 0, 0, st), n * 6)
    # generic fp32 -> fp32 group kernel (config 1 / weights)
    o32 = [torch.empty_like(t) for t in xf]
    if want("group e2m1 f32->f32"):
        res["group e2m1 f32->f32"] = timeit(lambda i: lib.fpq_fake_quant(xf[i].data_ptr(), o32[i].data_ptr(), n // 128, 128, 0, 0, 0, 0, 0, st), n * 8)
    if want("score 3 fmts f32 (read-only)"):
        res["score 3 fmts f32 (read-only)"] = timeit(lambda i: ops.score_formats(xf[i], ["e2m1", "e1m2", "e3m0"]), n * 4)
    if want("quant_grid f32 (quant_cuda.quant compat)"):
        grid = torch.tensor([-6, -4, -3, -2, -1.5, -1, -0.5, 0, 0.5, 1, 1.5, 2, 3, 4, 6.0], device=dev)
        res["quant_grid f32 (quant_cuda.quant compat)"] = timeit(lambda i: lib.fpq_quant_grid(xf[i].data_ptr(), grid.data_ptr(), 15, n, o32[i].data_ptr(), 0, st), n * 8)
    if want("weight transform+rotate f32 (fp64 butterflies)"):
        res["weight transform+rotate f32 (fp64 butterflies)"] = timeit(lambda i: lib.fpq_transform_rotate_weight(
            xf[i].data_ptr(), smooth.data_ptr(), sb, o32[i].data_ptr(), r4, C, st), n * 8)
    # plain copy for reference
    if want("torch copy f32 (d2d)"):
        res["torch copy f32 (d2d)"] = timeit(lambda i: o32[i].copy_(xf[i]), n * 8)
    for k, (g, us) in res.items():
        print(f"{k:32s} {g:8.1f} GB/s  {us:9.1f} us")


if __name__ == "__main__":
    main()
