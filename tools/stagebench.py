#!/usr/bin/env python
"""Per-stage and per-site breakdown of one bench step (development aid): one CUDA graph per (stage, site)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fpqvar_b200 import _lib as _L  # noqa: E402
if os.environ.get("FPQ_LIB_PATH"):
    _L.LIB_PATH = os.environ["FPQ_LIB_PATH"]
from fpqvar_b200.hotpath import DeviceReplay  # noqa: E402
from fpqvar_b200.var_workload import WORKLOADS  # noqa: E402

hot = WORKLOADS[os.environ.get("WORKLOAD", "var_d30_w4a4_rot")]
dev = torch.device("cuda")
C = hot.width
big = torch.randn(1 << 29, device=dev)                      # 2 GiB fp32
bigh = torch.nn.functional.gelu(torch.randn(1 << 30, device=dev)).half()   # 2 GiB fp16
outb = torch.empty(1 << 31, dtype=torch.uint8, device=dev)
smooth = {s: torch.exp(torch.rand(C, device=dev) * 2 - 1) for s in ("mat_qkv", "fc1")}
modulate = {s: (1.0 + 0.3 * torch.randn(2 * hot.batch, C, device=dev), 0.5 * torch.randn(2 * hot.batch, C, device=dev)) for s in ("mat_qkv", "fc1")}
rep = DeviceReplay(dev, smooth, modulate=modulate)
for kv in os.environ.get("FPQ_TUNABLES", "").split(","):        # e.g. FPQ_TUNABLES=rot_small_max_chunks=0,pdl=0
    if kv:
        k, v = kv.split("=")
        _L.set_tunable(k, int(v))
side = torch.cuda.Stream()
cur = {"f32": 0, "f16": 0, "out": 0}


def take(kind, nbytes, base, cap):
    n = (nbytes + 255) // 256 * 256
    if cur[kind] + n > cap:
        cur[kind] = 0
    p = base + cur[kind]
    cur[kind] += n
    return p


tot_t = tot_b = 0.0
print(f"{'stage':>5s} {'rows':>6s} {'site':>8s} {'MB/call':>9s} {'us/call':>9s} {'GB/s':>8s}")
for si, rows in enumerate(hot.stage_rows()):
    calls = [c for c in hot.calls() if c.stage == si]
    for site in ("mat_qkv", "proj", "fc1", "fc2"):
        sub = [c for c in calls if c.site == site]
        plan = []
        for c in sub:
            pin = take("f32", c.in_bytes, big.data_ptr(), 1 << 31) if c.in_dtype == "f32" else take("f16", c.in_bytes, bigh.data_ptr(), 1 << 31)
            plan.append((c, pin, take("out", c.out_bytes, outb.data_ptr(), 1 << 31)))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            for c, a, b in plan[:2]:
                rep.launch(c, a, b, side.cuda_stream)
            side.synchronize()
            with torch.cuda.graph(g, stream=side):
                for c, a, b in plan:
                    rep.launch(c, a, b, side.cuda_stream)
        for _ in range(2):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 1e3 / 5
        b = sum(c.bytes for c in sub)
        tot_t += t
        tot_b += b
        print(f"{si:5d} {rows:6d} {site:>8s} {b / len(sub) / 1e6:9.2f} {t / len(sub) * 1e6:9.2f} {b / t / 1e9:8.1f}")
print(f"sum: {tot_b / 1e9:.1f} GB in {tot_t * 1e3:.2f} ms -> {tot_b / tot_t / 1e9:.1f} GB/s")
