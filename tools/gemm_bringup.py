"""Bring-up and measurement of the low-bit GEMM on a B200 (tools; the parity tests proper are tests/test_gpu_gemm_codes.py).

  python tools/gemm_bringup.py            parity ladder against the oracle, then timings against torch's fp16 GEMM
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpqvar_b200 import _lib as L, lowbit, ops          # noqa: E402
from oracle import lowbit as LB, oracle as O             # noqa: E402  (checker only)


def bits(a):
    return a.view({1: np.uint8, 2: np.uint16, 4: np.uint32}[a.dtype.itemsize])


def same_bits(a, b):
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(bits(a)[~na], bits(b)[~nb])


def packed_from_oracle(q, s, fmt, dev):
    codes = torch.from_numpy(LB.to_blocked(LB.e4m3_encode(q))).to(dev)
    scales = torch.from_numpy(LB.scales_layout(s)).to(dev)
    return lowbit.PackedCodes(codes, scales, q.shape[0], q.shape[1], fmt)


def check_pack(dev):
    rng = np.random.default_rng(0)
    ok = True
    for dt in (np.float16, np.float32):
        for fmt in ("e2m1", "e1m2", "e3m0", "e2m3", "e3m2"):
            x = rng.standard_normal((300, 384)).astype(dt)
            x[0, :128] = 0; x[1, 5] = np.inf; x[2, 130] = np.nan
            p = lowbit.pack_codes(torch.from_numpy(x).to(dev), fmt)
            wc, ws = LB.pack_codes(x, fmt)
            gc, gs = p.codes.cpu().numpy(), p.scales.cpu().numpy()
            good = np.array_equal(gc, wc) and same_bits(gs, ws)
            out_dt = torch.float16 if dt == np.float16 or fmt in ("e2m3", "e3m2") else torch.float32
            dq = p.dequantize(out_dt).cpu().numpy()
            fq = O.fake_quant(x, fmt, 128, "kernel", out_dtype=dq.dtype)
            good2 = same_bits(dq, fq)
            if not (good and good2):
                ok = False
                print(f"  pack {fmt} {dt.__name__}: codes {np.array_equal(gc, wc)} scales {np.array_equal(bits(gs), bits(ws))} dequant {good2}"
                      f"  (code mismatches {(gc != wc).sum()}, scale mismatches {(bits(gs) != bits(ws)).sum()})")
    print("pack parity:", "OK" if ok else "FAILED")
    return ok


def gemm_case(dev, m, n, k, fmt="e2m1", seed=0, bias=False, out_dtype=torch.float32, pattern=None):
    rng = np.random.default_rng(seed)
    g = O.GRIDS[fmt].astype(np.float32)
    if pattern == "identity":
        qa = np.zeros((m, k), np.float32)
        qa[np.arange(m), np.arange(m) % k] = 1
        sa = np.ones((m, k // 128), np.float32)
        sw = np.ones((n, k // 128), np.float32)
    else:
        qa = rng.choice(g, (m, k)).astype(np.float32)
        sa = np.exp(rng.uniform(-2, 2, (m, k // 128))).astype(np.float32)
        sw = np.exp(rng.uniform(-4, 0, (n, k // 128))).astype(np.float32)
    qw = rng.choice(g, (n, k)).astype(np.float32)
    b = rng.standard_normal(n).astype(np.float32) if bias else None
    a = packed_from_oracle(qa, sa, fmt, dev)
    w = packed_from_oracle(qw, sw, fmt, dev)
    c = lowbit.linear_codes(a, w, None if b is None else torch.from_numpy(b).to(dev), out_dtype)
    torch.cuda.synchronize()
    want = LB.gemm_codes(qa, sa, qw, sw, b)
    got = c.cpu().numpy()
    if out_dtype == torch.float16:
        want = want.astype(np.float16)
    return got, want, (qa, sa, qw, sw)


def check_gemm(dev):
    ladder = [(128, 128, 128, None), (128, 128, 128, "identity"), (128, 128, 256, None), (128, 256, 384, None), (256, 128, 1920, None),
              (100, 136, 256, None), (300, 384, 7680, None)]
    all_ok = True
    for m, n, k, pat in ladder:
        got, want, ops_ = gemm_case(dev, m, n, k, pattern=pat)
        same = np.array_equal(bits(got), bits(want))
        all_ok &= same
        d = np.abs(got.astype(np.float64) - want.astype(np.float64))
        print(f"gemm {m}x{n}x{k} {pat or 'random'}: bit-exact {same}  max|diff| {np.nanmax(d):.4g}  max|want| {np.abs(want).max():.4g}  nan {np.isnan(got).sum()}")
        if not same and pat == "identity":
            qw = ops_[2]
            # C[i, j] should be qw[j, i]: report where the first rows actually come from
            for i in (0, 1, 8, 16, 17):
                col = got[i]
                match = [kk for kk in range(k) if np.array_equal(col, qw[:, kk])]
                print(f"   row {i}: equals weight column(s) {match[:4]}")
    return all_ok


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench(dev, shapes):
    torch.manual_seed(0)
    for name, m, n, k in shapes:
        x = torch.randn(m, k, device=dev, dtype=torch.float16)
        wt = (torch.randn(n, k, device=dev) * 0.02)
        a = lowbit.pack_codes(x, "e2m1")
        w = lowbit.pack_codes(wt, "e2m1")
        out = torch.empty(m, n, device=dev, dtype=torch.float16)
        xq = ops.fake_quant(x, "e2m1", 128, "kernel")
        wq = ops.fake_quant(wt, "e2m1", 128, "kernel").half()
        flop = 2.0 * m * n * k
        res = {}
        ar, wr = lowbit.pack_codes(x, "e2m1", True), lowbit.pack_codes(wt, "e2m1", True)
        for tn, ec, st, pair in ((256, 128, 6, 1), (256, 128, 4, 0), (256, 64, 4, 0), (128, 32, 6, 0), (128, 64, 6, 0), (128, 128, 6, 0)):
            L.set_tunable("gemm_stages", st)
            L.set_tunable("gemm_tile_n", tn)
            L.set_tunable("gemm_epi_cols", ec)
            L.set_tunable("gemm_pair", pair)
            tag = f"tn{tn} ec{ec}" + (" pair" if pair else "")
            res[f"g128 {tag}"] = timeit(lambda: lowbit.linear_codes(a, w, None, torch.float16, out))
            res[f"row {tag}"] = timeit(lambda: lowbit.linear_codes(ar, wr, None, torch.float16, out))
        L.set_tunable("gemm_stages", 6)
        L.set_tunable("gemm_tile_n", 256)
        L.set_tunable("gemm_epi_cols", 128)
        L.set_tunable("gemm_pair", -1)
        res["sse g128"] = timeit(lambda: lowbit.linear_codes_sse(a, w, out))
        t_packrow = timeit(lambda: lowbit.pack_codes(x, "e2m1", True))
        t_pack = timeit(lambda: lowbit.pack_codes(x, "e2m1"))
        t_fq = timeit(lambda: ops.fake_quant(x, "e2m1", 128, "kernel"))
        t_ref = timeit(lambda: torch.nn.functional.linear(xq, wq))
        y = lowbit.linear_codes(a, w, None, torch.float16)
        yr = torch.nn.functional.linear(xq, wq)
        rel = ((y.float() - yr.float()).abs().max() / yr.float().abs().max()).item()
        print(f"{name} m={m} n={n} k={k}: codes GEMM " + " ".join(f"[{st}] {t:.3f} ms ({flop / t / 1e9:.0f} TF/s)" for st, t in res.items())
              + f" | cuBLAS fp16 on fake-quantized {t_ref:.3f} ms ({flop / t_ref / 1e9:.0f} TF/s) | pack {t_pack:.3f} ms ({(m * k * 3 + m * k / 32) / t_pack / 1e6:.0f} GB/s)"
              f" pack per-row {t_packrow:.3f} ms fake_quant {t_fq:.3f} ms | max rel diff {rel:.2e}", flush=True)


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    print(torch.cuda.get_device_name(0))
    t0 = time.time()
    ok = check_pack(dev)
    L.set_tunable("gemm_pair", 1)
    print("CTA pairs:")
    ok2 = check_gemm(dev)
    L.set_tunable("gemm_pair", 0)
    print("single CTAs:")
    ok2 = check_gemm(dev) and ok2
    for tn, ec in ((256, 64), (128, 32), (128, 64), (128, 128)):
        L.set_tunable("gemm_tile_n", tn)
        L.set_tunable("gemm_epi_cols", ec)
        print(f"tile_n {tn}, epilogue columns {ec}:")
        ok2 = check_gemm(dev) and ok2
    L.set_tunable("gemm_tile_n", 256)
    L.set_tunable("gemm_epi_cols", 128)
    L.set_tunable("gemm_pair", -1)
    print(f"parity ladder took {time.time() - t0:.1f} s")
    if ok2 or "--bench" in sys.argv:
        shapes = [("d30 mat_qkv stage 9", 25600, 5760, 1920), ("d30 fc1 stage 9", 25600, 7680, 1920), ("d30 proj stage 9", 25600, 1920, 1920),
                  ("d30 fc2-shape stage 9", 25600, 1920, 7680), ("d30 mat_qkv all stages", 68000, 5760, 1920), ("d16 fc1 B=64", 32768, 4096, 1024)]
        bench(dev, shapes[:1] + shapes[3:4] if "--short" in sys.argv else shapes)
