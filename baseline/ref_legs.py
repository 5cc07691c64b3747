"""MEASUREMENT INFRASTRUCTURE ONLY: the reference's OWN code timed on the bench workload (bench.py's reference legs).

  cpu_reference_sample   the reference's torch CPU path -- the non-`_cuda` functions of
                         models_fp_quant_transform_rotate/quant_utils.py (fp_quant_e2_per_group :298,
                         fp_quant_e1m2_neg_e2m1_pos_per_group :381, fp6 via quantize_to_nearest_grid) and, at mat_qkv / fc1,
                         the reference's online sequence basic_var.py:263,266 in front of them -- on a BOUNDED sample of the
                         step, on the box's host cores (torch's intra-op threads)
  gpu_reference_step     the reference's GPU path -- the same call sites with its fp_quant_*_cuda functions around its own
                         quant_cuda extension (compiled unmodified for sm_100a) -- over every call of one step, CUDA events

Both import the unmodified reference through baseline/ref_env.py; nothing of this repository's library runs here.
"""
from __future__ import annotations

import os
import time


def _cpu_quantizers(ns, hot):
    """(activation quantizer, fc2 quantizer) of the workload as the reference's CPU functions.  FP4: its non-`_cuda`
    functions (torch.argmin rounding).  FP6 has no non-`_cuda` function in the reference; there its `_cuda` function runs
    with `quant_cuda.quant` (the only CUDA-only line, qu.py:513) answered by the reference's own CPU statement of the same
    rounding, `quantize_to_nearest_grid` (qu.py:209-230) -- every other line is the reference's glue, unmodified."""
    qu = ns.qu

    class _CpuQuant:
        @staticmethod
        def quant(x, grid):
            return qu.quantize_to_nearest_grid(x, grid.to(x.dtype)), None

    def on_cpu(fn):
        def run(x, n_bits, group_size=128):
            saved = qu.quant_cuda
            qu.quant_cuda = _CpuQuant
            try:
                return fn(x, n_bits, group_size)
            finally:
                qu.quant_cuda = saved
        return run

    act = {"e2m1": qu.fp_quant_e2_per_group, "e1m2": qu.fp_quant_e1_per_group, "e3m0": qu.fp_quant_e3_per_group,
           "e2m3": on_cpu(qu.fp6_quant_e2m3_per_group_cuda), "e3m2": on_cpu(qu.fp6_quant_e3m2_per_group_cuda)}[hot.act_fmt]
    if hot.fc2_op != "signsplit":
        return act, act
    fc2 = {"e1m2_neg_e2m1_pos": qu.fp_quant_e1m2_neg_e2m1_pos_per_group,
           "int_neg_e2m3_pos": on_cpu(qu.fp6_quant_int_neg_e2m3_pos_per_group_cuda)}[hot.fc2_fmt]
    return act, fc2


def cpu_sample_calls(hot, max_rows):
    """The bounded sample: the four quantizer calls of ONE block at every stage, token rows capped at `max_rows` per call."""
    import dataclasses
    out = []
    for c in hot.calls(blocks=[0]):
        rows = min(c.rows, max_rows)
        if c.rows_per_batch:
            rows = max(c.rows_per_batch, rows // c.rows_per_batch * c.rows_per_batch)
        out.append(dataclasses.replace(c, rows=rows))
    return out


def cpu_reference_sample(hot, max_rows=2048, steps=1, warmup=1, budget_s=None, bits=4):
    """Times the reference's torch CPU path.  Returns (dict for the JSON line, mean seconds per step, steps)."""
    import torch
    from . import ref_env
    ns = ref_env.load("cpu")
    act, fc2 = _cpu_quantizers(ns, hot)
    C = hot.width
    g = torch.Generator().manual_seed(0)
    Q = ns.rotation_utils.block_random_hadamard_matrix(total_size=C, block_size=128, device="cpu", seed=42).to(torch.float32)
    s = torch.exp(torch.rand(C, generator=g) * 2 - 1)
    n_bits = 6 if hot.act_fmt in ("e2m3", "e3m2") else 4

    def make(calls):
        data = []
        for c in calls:
            x = torch.randn(c.rows, c.cols, generator=g)
            if c.site == "fc2":
                x = torch.nn.functional.gelu(x, approximate="tanh")
            if c.in_dtype == "f16":
                x = x.half()
            extra = None
            if c.op == "mod_rotate_quant":
                b = c.rows // c.rows_per_batch
                extra = (0.3 * torch.randn(b, 1, C, generator=g), 0.5 * torch.randn(b, 1, C, generator=g))
            data.append((x, extra))
        return data

    def run(calls, data):
        with torch.no_grad():
            for c, (x, extra) in zip(calls, data):
                if c.op in ("rotate_quant", "mod_rotate_quant"):
                    t = x
                    if extra is not None:                                     # basic_var.py:263: .mul(scale.add(1)).add_(shift)
                        b = c.rows // c.rows_per_batch
                        t = t.view(b, c.rows_per_batch, C).mul(extra[0].add(1)).add_(extra[1]).view(c.rows, C)
                    t = torch.matmul(t.mul(s), Q).half()                      # .mul(s) @ Q; the autocast GEMM returns fp16
                    act(t, n_bits, 128)
                elif c.op == "signsplit":
                    fc2(x.clone(), n_bits, 128)
                else:
                    act(x.clone(), n_bits, 128)                               # fp_quant_e2_per_group mutates its argument (qu.py:306)

    calls = cpu_sample_calls(hot, max_rows)
    data = make(calls)
    t0 = time.perf_counter()
    run(calls, data)                                                          # warm-up 1 (also sizes the sample)
    t1 = time.perf_counter() - t0
    if budget_s is not None and t1 * (steps + warmup) > budget_s and max_rows > 256:
        max_rows = max(256, int(max_rows * budget_s / (t1 * (steps + warmup))) // 64 * 64)
        calls = cpu_sample_calls(hot, max_rows)
        data = make(calls)
        run(calls, data)
    for _ in range(max(0, warmup - 1)):
        run(calls, data)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        run(calls, data)
        times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    nbytes = sum(c.bytes for c in calls)
    threads = torch.get_num_threads()
    return {
        "value": nbytes / mean / 1e9, "unit": "GB/s", "cores": threads, "kind": "reference",
        "host_cpus": os.cpu_count(),
        "sample": f"the reference's torch CPU functions (baseline/_ref: fp_quant_*_per_group + the basic_var.py:263 sequence) on 1 of {hot.depth} "
                  f"blocks x all {len(hot.patch_nums)} stages of {hot.name}, token rows capped at {max_rows} per call "
                  f"({nbytes / 1e9:.3f} GB algorithmic, {len(calls)} calls), {len(times)} timed passes of {mean:.2f} s, torch threads = {threads}",
    }, mean, len(times)


def gpu_reference_step(hot, dev, iters=2):
    """The reference's GPU path over every call of one step (its own Python glue + its own extension), CUDA events."""
    import torch
    from . import ref_env
    ns = ref_env.load(str(dev))
    qu = ns.qu
    act = {"e2m1": qu.fp_quant_e2_per_group_cuda, "e1m2": qu.fp_quant_e1_per_group_cuda, "e3m0": qu.fp_quant_e3_per_group_cuda,
           "e2m3": qu.fp6_quant_e2m3_per_group_cuda, "e3m2": qu.fp6_quant_e3m2_per_group_cuda}[hot.act_fmt]
    fc2 = {"e1m2_neg_e2m1_pos": qu.fp_quant_e1m2_neg_e2m1_pos_per_group_cuda,
           "int_neg_e2m3_pos": qu.fp6_quant_int_neg_e2m3_pos_per_group_cuda}.get(hot.fc2_fmt, act) if hot.fc2_op == "signsplit" else act
    n_bits = 6 if hot.act_fmt in ("e2m3", "e3m2") else 4
    C = hot.width
    g = torch.Generator(device=dev).manual_seed(5)
    rows_max = max(hot.stage_rows())
    x32 = torch.randn(rows_max, C, device=dev, generator=g)
    x16 = torch.randn(rows_max, C, device=dev, generator=g).half()
    xg = torch.nn.functional.gelu(torch.randn(rows_max, 4 * C, device=dev, generator=g), approximate="tanh").half()
    Q = ns.rotation_utils.block_random_hadamard_matrix(total_size=C, block_size=128, device=dev, seed=42).to(torch.float32)
    s = torch.exp(torch.rand(C, device=dev, generator=g) * 2 - 1)
    scale = (0.3 * torch.randn(2 * hot.batch, 1, C, device=dev, generator=g)).half()
    shift = (0.5 * torch.randn(2 * hot.batch, 1, C, device=dev, generator=g)).half()
    calls = hot.calls()
    per_site = {}

    def one(c):
        if c.op in ("rotate_quant", "mod_rotate_quant"):
            t = x32[:c.rows]
            if c.op == "mod_rotate_quant":
                b = c.rows // c.rows_per_batch
                t = t.view(b, c.rows_per_batch, C).mul(scale[:b].add(1)).add_(shift[:b])       # basic_var.py:263
            t = torch.matmul(t.mul(s), Q)                                                     # fp16 under autocast
            act(t.view(c.rows, C), n_bits, 128)
        elif c.op == "signsplit":
            fc2(xg[:c.rows], n_bits, 128)
        else:
            src = x16 if c.in_dtype == "f16" else x32
            act(src[:c.rows], n_bits, 128)

    def run(subset):
        with torch.inference_mode():
            with torch.autocast("cuda", enabled=True, dtype=torch.float16, cache_enabled=True):    # evaluate_fp_quant_transform_rotate.py:195
                for c in subset:
                    one(c)

    run([c for c in calls if c.block == 0])                                   # warm-up
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run(calls)
    e1.record()
    torch.cuda.synchronize(dev)
    t_step = e0.elapsed_time(e1) / 1e3 / iters
    # the largest stage of every site alone
    last = max(c.stage for c in calls)
    for site in ("mat_qkv", "proj", "fc1", "fc2"):
        sub = [c for c in calls if c.stage == last and c.site == site][:4]
        run(sub[:1])
        torch.cuda.synchronize(dev)
        e0.record()
        run(sub)
        e1.record()
        torch.cuda.synchronize(dev)
        per_site[site] = sum(c.bytes for c in sub) / (e0.elapsed_time(e1) / 1e3) / 1e9
    nbytes = sum(c.bytes for c in calls)
    return {"GB/s": nbytes / t_step / 1e9, "ms_per_step": t_step * 1e3, "iters": iters, "calls_per_step": len(calls),
            "largest_stage_GB/s_by_site": per_site,
            "what": "the reference's own online sequence (basic_var.py:263,266 modulate, .mul(s), dense [C,C] rotation GEMM under fp16 autocast) "
                    "and fp_quant_*_cuda functions (baseline/_ref, unmodified) around its quant_cuda extension compiled for sm_100a, "
                    "same calls and algorithmic bytes as the step above"}


if __name__ == "__main__":          # child-process entry used by bench.py: python -m baseline.ref_legs gpu_step <workload>
    import json
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    from fpqvar_b200.var_workload import WORKLOADS
    if len(sys.argv) >= 3 and sys.argv[1] == "gpu_step":
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        torch.cuda.set_device(dev)
        print(json.dumps(gpu_reference_step(WORKLOADS[sys.argv[2]], dev, iters=2)), flush=True)
