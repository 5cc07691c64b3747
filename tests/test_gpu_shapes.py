"""GPU: shape / alignment / layout coverage of the C-ABI entry points against the oracle -- group sizes other
than 128, rows of odd lengths, views that are only 2- or 4-byte aligned, non-contiguous inputs, every dtype
pair of the row kernels, and use from a side stream with and without programmatic dependent launch."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import bits_equal, mismatch_report
from oracle import oracle as O

pytestmark = pytest.mark.gpu
NP = {torch.float32: np.float32, torch.float16: np.float16}


@pytest.fixture(scope="module")
def ops():
    from fpqvar_b200 import ops as _ops
    return _ops


def host(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("group", [16, 32, 64, 128, 256, 512, 1920])
@pytest.mark.parametrize("dt", [torch.float32, torch.float16])
def test_group_sizes(ops, group, dt):
    torch.manual_seed(group)
    x = (torch.randn(77, group * 3, device="cuda") * 1.7).to(dt)
    for fmt, tie in (("e2m1", "kernel"), ("e3m2", "kernel"), ("e1m2", "argmin")):
        want = O.fake_quant(host(x), fmt, group, tie)
        got = host(ops.fake_quant(x, fmt, group, tie))
        assert bits_equal(got, want), f"group {group} {fmt} {tie}\n" + mismatch_report(got, want)
    want = O.fake_quant_signsplit(host(x), "e1m2_neg_e2m1_pos", group, "kernel", clipping_strength=None)
    got = host(ops.fake_quant_signsplit(x, "e1m2_neg_e2m1_pos", group, "kernel"))
    assert bits_equal(got, want), f"split group {group}"


@pytest.mark.parametrize("row_len", [256, 264, 1000, 1920, 2304, 4096, 7680, 9216, 16384, 40000])
@pytest.mark.parametrize("din,dout", [(torch.float16, torch.float16), (torch.float32, torch.float16), (torch.float32, torch.float32),
                                      (torch.float16, torch.float32)])
def test_row_kernels_every_dtype_pair(ops, row_len, din, dout):
    torch.manual_seed(row_len)
    x = (torch.randn(19, row_len, device="cuda") * torch.exp(torch.randn(19, 1, device="cuda"))).to(din)
    x[3] = 0
    x[5, 7] = float("nan")
    x[6, 1] = float("inf")
    tie = "argmin" if dout == torch.float32 and din == torch.float16 else "kernel"
    for fmt in ("e2m3", "e2m1"):
        want = O.fake_quant(host(x), fmt, None, tie, clamp3=(tie == "argmin"), out_dtype=NP[dout])
        got = host(ops.fake_quant(x, fmt, None, tie, clamp3=(tie == "argmin"), out_dtype=dout))
        assert bits_equal(got, want), f"rows {row_len} {din}->{dout} {fmt}\n" + mismatch_report(got, want)


@pytest.mark.parametrize("row_len", [256, 264, 1000, 1920, 2304, 4096, 7680, 9216, 16384, 40000])
@pytest.mark.parametrize("din", [torch.float16, torch.float32])
@pytest.mark.parametrize("tie", ["kernel", "argmin"])
def test_signsplit_row_kernels(ops, row_len, din, tie):
    """Per-token sign-split (fp6_quant_int_neg_e2m3_pos_per_token_cuda qu.py:614-646 and friends) through the
    row-in-registers kernel (every values-per-thread choice) and the two-pass fallback: one-sided rows (a zero scale on
    the other side), zero / NaN / +-inf rows, rows whose scales are fp16-subnormal."""
    torch.manual_seed(row_len + 1)
    x = (torch.randn(23, row_len, device="cuda") * torch.exp(torch.randn(23, 1, device="cuda") * 2)).to(din)
    x[3] = 0
    x[4] = x[4].abs()                                        # no negative element: sn = 0
    x[5] = -x[5].abs()                                       # no positive element: sp = 0
    x[6, 7] = float("nan")
    x[7, 1] = float("inf")
    x[8, row_len - 1] = float("-inf")
    x[9] = x[9] * 1e-6                                       # scales in the fp16 subnormal range
    x[10] = torch.where(x[10] > 0, x[10], x[10] * 0.02)      # GELU-like skew
    for fmt in ("int_neg_e2m3_pos", "e1m2_neg_e2m1_pos"):
        want = O.fake_quant_signsplit(host(x), fmt, None, tie, clipping_strength=None)
        got = host(ops.fake_quant_signsplit(x, fmt, None, tie))
        assert bits_equal(got, want), f"split rows {row_len} {din} {fmt} {tie}\n" + mismatch_report(got, want)


@pytest.mark.parametrize("v", [1, 2, 4])
def test_row_kernels_values_per_thread_override(ops, v):
    """The tunable row_v (values per thread of the per-token kernels, a measurement aid) must not change results."""
    from fpqvar_b200 import _lib as L
    L.set_tunable("row_v", v)
    try:
        torch.manual_seed(v)
        for row_len in (520, 1920, 7680):
            x = (torch.randn(9, row_len, device="cuda") * 3).half()
            x[2, 5] = float("nan")
            assert bits_equal(host(ops.fake_quant(x, "e2m3", None, "kernel")), O.fake_quant(host(x), "e2m3", None, "kernel"))
            assert bits_equal(host(ops.fake_quant_signsplit(x, "int_neg_e2m3_pos", None, "kernel")),
                              O.fake_quant_signsplit(host(x), "int_neg_e2m3_pos", None, "kernel", clipping_strength=None))
    finally:
        L.set_tunable("row_v", 0)


@pytest.mark.parametrize("dt", [torch.float32, torch.float16])
def test_unaligned_and_noncontiguous_inputs(ops, dt):
    torch.manual_seed(1)
    base = torch.randn(128 * 50 + 3, device="cuda").to(dt)
    for off in (1, 2, 3):                                   # 2/4-byte aligned views: the vector paths must not be taken
        v = base[off: off + 128 * 50]
        want = O.fake_quant(host(v), "e2m1", 128, "kernel")
        assert bits_equal(host(ops.fake_quant(v, "e2m1", 128, "kernel")), want), f"offset {off}"
        want = O.fake_quant_signsplit(host(v), "e1m2_neg_e2m1_pos", 128, "kernel", clipping_strength=None)
        assert bits_equal(host(ops.fake_quant_signsplit(v, "e1m2_neg_e2m1_pos", 128, "kernel")), want), f"split offset {off}"
    t = torch.randn(64, 512, device="cuda").to(dt).t()      # non-contiguous: made contiguous by the host layer
    want = O.fake_quant(np.ascontiguousarray(host(t)), "e2m1", 128, "kernel")
    assert bits_equal(host(ops.fake_quant(t, "e2m1", 128, "kernel")), want)
    rows = torch.randn(33, 1000, device="cuda").to(dt)[:, 1:]      # rows of 999 elements starting at odd offsets
    want = O.fake_quant(host(rows), "e2m3", None, "kernel", out_dtype=np.float16)
    assert bits_equal(host(ops.fake_quant(rows, "e2m3", None, "kernel", out_dtype=torch.float16)), want)


def test_side_stream_and_graph_replay(ops):
    torch.manual_seed(2)
    x = torch.nn.functional.gelu(torch.randn(4096, 1024, device="cuda")).half()
    want = O.fake_quant_signsplit(host(x), "e1m2_neg_e2m1_pos", 128, "kernel")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        y = ops.fake_quant_signsplit(x, "e1m2_neg_e2m1_pos", 128, "kernel", global_clip=True)
        z = ops.fake_quant(y, "e2m1", 128, "kernel")                      # dependent launch right behind (PDL edge)
    s.synchronize()
    assert bits_equal(host(y), want)
    assert bits_equal(host(z), O.fake_quant(want, "e2m1", 128, "kernel"))
    # NaN poisoning through the whole-tensor clip, then a clean call on the same stream/workspace
    xn = x.clone()
    xn[17, 5] = float("nan")
    with torch.cuda.stream(s):
        yn = ops.fake_quant_signsplit(xn, "e1m2_neg_e2m1_pos", 128, "kernel", global_clip=True)
        y2 = ops.fake_quant_signsplit(x, "e1m2_neg_e2m1_pos", 128, "kernel", global_clip=True)
    s.synchronize()
    assert bits_equal(host(yn), O.fake_quant_signsplit(host(xn), "e1m2_neg_e2m1_pos", 128, "kernel"))
    assert float(yn.abs().max()) == 0.0
    assert bits_equal(host(y2), want), "workspace not reset after a poisoned call"


@pytest.mark.parametrize("kb", [64, 100, 228])
def test_results_identical_under_a_fixed_shared_memory_carveout(ops, kb):
    """The tunable smem_kb (one L1 / shared-memory split for every activation kernel instead of the driver's choice; it also
    bounds the streaming rotate kernel's CTAs per SM) is a performance knob only: same bits as the default."""
    from fpqvar_b200 import _lib as L
    from fpqvar_b200.hotpath import seed42_sign_bits
    torch.manual_seed(3)
    x = torch.randn(6000, 1920, device="cuda")                      # 90 000 chunks: the streaming kernel
    h = torch.nn.functional.gelu(x).half()
    s_ = torch.exp(torch.rand(1920, device="cuda") * 2 - 1)
    sc = torch.randn(60, 1, 1920, device="cuda") * 0.3
    sh = torch.randn(60, 1, 1920, device="cuda") * 0.5

    def run():
        a = ops.transform_rotate_quant(x, s_, seed42_sign_bits(), "e2m1")
        m = ops.modulate_transform_rotate_quant(x.view(60, 100, 1920), sc, sh, s_, seed42_sign_bits(), "e2m1")
        b = ops.fake_quant_signsplit(h, "e1m2_neg_e2m1_pos", 128, "kernel", global_clip=True)
        c = ops.fake_quant(h, "e2m1", 128, "kernel")
        torch.cuda.synchronize()
        return [host(t) for t in (a, m, b, c)]

    default = run()
    L.set_tunable("smem_kb", kb)
    try:
        fixed = run()
    finally:
        L.set_tunable("smem_kb", 0)
    for u, v in zip(default, fixed):
        assert bits_equal(u, v)


def test_results_identical_without_pdl(ops):
    """Plain stream-ordered launches (tunable pdl = 0) must give the same bits as programmatic dependent launches."""
    from fpqvar_b200 import _lib as L
    from fpqvar_b200.hotpath import seed42_sign_bits
    torch.manual_seed(0)
    x = torch.randn(3000, 1920, device="cuda")
    h = torch.nn.functional.gelu(x).half()
    s_ = torch.exp(torch.rand(1920, device="cuda") * 2 - 1)

    def run():
        a = ops.transform_rotate_quant(x, s_, seed42_sign_bits(), "e2m1")
        b = ops.fake_quant_signsplit(h, "e1m2_neg_e2m1_pos", 128, "kernel", global_clip=True)
        c = ops.fake_quant(h, "e2m1", 128, "kernel")
        torch.cuda.synchronize()
        return [host(t) for t in (a, b, c)]

    with_pdl = run()
    L.set_tunable("pdl", 0)
    try:
        without = run()
    finally:
        L.set_tunable("pdl", 1)
    for u, v in zip(with_pdl, without):
        assert bits_equal(u, v)


def test_rotate_kernel_choice_does_not_change_results(ops):
    """The tunable rot_small_max_chunks moves the boundary between the small-launch rotate kernel and the streaming
    (bulk-copy staged) one: the same tensor must rotate and quantize to the same bits through either, with and without the
    adaLN modulate, for every row length the streaming kernel plans differently (chunk columns per warp 1 / 2 / 4, idle lane
    sets, 8 or 9 consumer warps)."""
    from fpqvar_b200 import _lib as L
    from fpqvar_b200.hotpath import seed42_sign_bits
    sb = seed42_sign_bits()
    torch.manual_seed(11)
    try:
        for cols, rows, rpb in ((128, 777, 7), (256, 500, 50), (384, 301, 43), (640, 257, 257), (1024, 130, 65), (1920, 143, 11),
                                (2304, 97, 97), (2944, 50, 25), (4608, 37, 37)):
            x = torch.randn(rows, cols, device="cuda") * torch.exp(torch.randn(rows, 1, device="cuda"))
            x[rows // 2, : 128] = 0.0                                   # an all-zero group (irregular scale)
            s_ = torch.exp(torch.rand(cols, device="cuda") * 2 - 1)
            b = rows // rpb
            sc = torch.randn(b, 1, cols, device="cuda") * 0.3
            sh = torch.randn(b, 1, cols, device="cuda") * 0.5
            res = []
            for small_max in (1 << 40, 0):                              # everything small / everything streaming
                L.set_tunable("rot_small_max_chunks", small_max)
                q, r = ops.transform_rotate_quant(x, s_, sb, "e2m1", return_rotated=True)
                q6 = ops.transform_rotate_quant(x, None, sb, "e3m2")
                qm, rm = ops.modulate_transform_rotate_quant(x.view(b, rpb, cols), sc, sh, s_, sb, "e2m3", return_rotated=True)
                res.append([host(t) for t in (q, r, q6, qm, rm)])
            for u, v in zip(*res):
                assert bits_equal(u, v), (cols, rows, rpb)
    finally:
        L.set_tunable("rot_small_max_chunks", 40000)


def test_randomized_differential_against_oracle(ops):
    """60 random (shape, dtype, format, tie, scale distribution) cases against the oracle -- a cheap fuzz of the dispatch
    (group kernels, packed fp16 path, row-in-registers kernels, two-pass fallback)."""
    rng = np.random.default_rng(20261018)
    syms = ["e2m1", "e1m2", "e3m0", "e2m3", "e3m2"]
    splits = ["e1m2_neg_e2m1_pos", "int_neg_e2m3_pos", "afpq_e2m1"]
    for case in range(60):
        dt = [torch.float16, torch.float32][int(rng.integers(2))]
        per_row = bool(rng.integers(2))
        if per_row:
            row_len = int(rng.choice([64, 96, 128, 200, 256, 520, 1024, 1920, 3000, 4608, 8192, 9216, 20000]))
            rows = int(rng.integers(1, 40))
            shape, group = (rows, row_len), None
        else:
            group = int(rng.choice([32, 64, 128, 128, 128, 256]))
            shape, _ = (int(rng.integers(1, 300)), group * int(rng.integers(1, 20))), None
        x = rng.standard_normal(shape).astype(np.float32) * np.exp(rng.uniform(-6, 4, (shape[0], 1))).astype(np.float32)
        kind = int(rng.integers(4))
        if kind == 0:
            x = np.where(x > 0, x, x * 0.03).astype(np.float32)          # GELU-like skew
        elif kind == 1:
            x[rng.integers(shape[0])] = 0.0
        elif kind == 2 and x.size > 10:
            x.reshape(-1)[rng.integers(x.size)] = [np.nan, np.inf, -np.inf][int(rng.integers(3))]
        with np.errstate(over="ignore"):
            xh = x.astype(NP[dt])
        xt = torch.from_numpy(xh).cuda()
        tie = ["kernel", "argmin"][int(rng.integers(2))]
        if rng.integers(3) == 0:
            fmt = splits[int(rng.integers(3))]
            want = O.fake_quant_signsplit(xh, fmt, group, tie, clipping_strength=None)
            got = host(ops.fake_quant_signsplit(xt, fmt, group, tie))
        else:
            fmt = syms[int(rng.integers(5))]
            clamp3 = bool(rng.integers(2)) and tie == "argmin"
            want = O.fake_quant(xh, fmt, group, tie, clamp3=clamp3)
            got = host(ops.fake_quant(xt, fmt, group, tie, clamp3=clamp3))
        assert bits_equal(got, want), f"case {case}: {shape} {dt} group={group} {fmt} {tie}\n" + mismatch_report(got, want)


def test_concurrent_host_threads_and_streams(ops):
    """The C ABI is called from several host threads at once (ctypes drops the GIL), each on its own stream: per-thread
    error / launch bookkeeping, per-(device, stream) clip workspaces, benign static caches.  Results must equal the
    serial ones and every thread must see its own launch count."""
    import threading
    from fpqvar_b200.hotpath import seed42_sign_bits
    bits = seed42_sign_bits()
    n_threads, iters = 4, 40
    torch.manual_seed(5)
    xs = [torch.randn(300 + 7 * t, 1920, device="cuda") for t in range(n_threads)]
    hs = [torch.nn.functional.gelu(x).half() for x in xs]
    s = torch.exp(torch.rand(1920, device="cuda") - 0.5)
    want = [(ops.transform_rotate_quant(x, s, bits, "e2m1"), ops.fake_quant_signsplit(h, "e1m2_neg_e2m1_pos", 128, "kernel", global_clip=True),
             ops.fake_quant(h, "e2m3", None, "kernel")) for x, h in zip(xs, hs)]
    torch.cuda.synchronize()
    errors, counts = [], [0] * n_threads

    def worker(t):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                n0 = ops.launch_count()
                for _ in range(iters):
                    got = (ops.transform_rotate_quant(xs[t], s, bits, "e2m1"),
                           ops.fake_quant_signsplit(hs[t], "e1m2_neg_e2m1_pos", 128, "kernel", global_clip=True),
                           ops.fake_quant(hs[t], "e2m3", None, "kernel"))
                st.synchronize()
                counts[t] = ops.launch_count() - n0
                for g, w in zip(got, want[t]):
                    if not torch.equal(g.view(torch.int16), w.view(torch.int16)):
                        errors.append(f"thread {t}: result differs")
        except Exception as e:  # noqa: BLE001
            errors.append(f"thread {t}: {type(e).__name__}: {e}")

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    assert counts == [3 * iters] * n_threads


def test_whole_tensor_clip_protocol_under_stress(ops):
    """The {flag, ticket} protocol of the packed sign-split kernel (fire-and-forget wrapping tickets, one owner CTA for the
    rewrite): clean and poisoned calls of different sizes back to back on one stream and workspace, NaNs in one group, in
    many groups and in EVERY group (every CTA sees one: still exactly one owner), single-CTA grids -- every output is what
    the reference's clamp gives, and the workspace is zero afterwards."""
    from fpqvar_b200 import ops as OPS
    g = torch.Generator(device="cuda").manual_seed(77)
    ws = OPS._clip_workspace(torch.device("cuda", torch.cuda.current_device()))
    sizes = [(1, 128), (3, 256), (100, 7680), (4096, 1024), (25600, 1920), (7, 128 * 9)]
    results = []
    for rnd in range(3):
        for rows, cols in sizes:
            x = torch.nn.functional.gelu(torch.randn(rows, cols, device="cuda", generator=g)).half()
            kind = (rnd + rows) % 4
            if kind == 1:
                x[rows // 2, 3] = float("nan")
            elif kind == 2:
                x[:, ::128] = float("nan")                             # a NaN in every group
            elif kind == 3:
                x[torch.rand(rows, cols, device="cuda", generator=g) < 1e-4] = float("nan")
            y = ops.fake_quant_signsplit(x, "e1m2_neg_e2m1_pos", 128, "kernel", global_clip=True)
            results.append((x, y, bool(torch.isnan(x).any())))
    torch.cuda.synchronize()
    assert ws.tolist() == [0, 0], "workspace not left zeroed"
    for x, y, poisoned in results:
        if poisoned:
            assert float(y.abs().max()) == 0.0 and not bool(torch.isnan(y).any())
        else:
            assert bits_equal(host(y), O.fake_quant_signsplit(host(x), "e1m2_neg_e2m1_pos", 128, "kernel"))
