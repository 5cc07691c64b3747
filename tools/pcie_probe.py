#!/usr/bin/env python
"""Raw pinned-memory PCIe rates on this box (development aid): H2D alone, D2H alone, both at once -- on one GPU, or on all
ranks at once under torchrun (what bounds bench.py's e2e leg at N > 1: every rank streams 70 GB up and 55 GB down per step).

    python tools/pcie_probe.py
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py [--bind]

--bind pins the process to the CPUs of its GPU's NUMA node BEFORE the pinned buffers are allocated (first touch places them
on that node), the same thing bench.py does for its e2e leg (fpqvar_b200.hotpath.bind_to_gpu_numa)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpqvar_b200.hotpath import bind_to_gpu_numa, gpu_numa_node  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
bound = bind_to_gpu_numa(local) if "--bind" in sys.argv else None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
h_out.fill_(1)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=8):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * n / dt / 1e9


run(True, True, 2)
res = {"rank": rank, "gpu_numa_node": gpu_numa_node(local), "bound_cpus": bound, "cpus_allowed": len(os.sched_getaffinity(0)),
       "h2d_alone": round(run(True, False), 1), "d2h_alone": round(run(False, True), 1), "both_each_way": round(run(True, True), 1)}
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, res)
    if rank == 0:
        for r in out:
            print(json.dumps(r))
        print(json.dumps({"ranks": world, "aggregate_h2d_alone": round(sum(r["h2d_alone"] for r in out), 1),
                          "aggregate_d2h_alone": round(sum(r["d2h_alone"] for r in out), 1),
                          "aggregate_both_each_way": round(sum(r["both_each_way"] for r in out), 1), "unit": "GB/s, all ranks at once"}))
    dist.destroy_process_group()
else:
    print(json.dumps(res))
