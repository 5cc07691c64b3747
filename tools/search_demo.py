#!/usr/bin/env python
"""BASELINE configs[4] in miniature: FP4 (or FP6) format search over synthetic calibration activations with the
reference's shapes, (layer, weight-format) units sharded over the ranks, one final all-reduce (NCCL) of the loss table.

    python tools/search_demo.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/search_demo.py
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpqvar_b200 import search  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    C = int(os.environ.get("SEARCH_C", "1920"))
    n_blocks = int(os.environ.get("SEARCH_BLOCKS", "4"))
    n_act = int(os.environ.get("SEARCH_ACTS", "20"))
    patch = (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)
    g = torch.Generator(device=dev).manual_seed(0)              # same data on every rank (the reference reads the same files)
    layers = []
    for b in range(n_blocks):
        for name, (o, i, dt, gelu) in {"mat_qkv": (3 * C, C, torch.float32, False), "proj": (C, C, torch.float16, False),
                                       "fc1": (4 * C, C, torch.float32, False), "fc2": (C, 4 * C, torch.float16, True)}.items():
            w = (torch.randn(o, i, device=dev, generator=g) * 0.02).to(dt)
            acts = []
            for j in range(n_act):
                x = torch.randn(2, patch[j % 10] ** 2, i, device=dev, generator=g)
                if gelu:
                    x = torch.nn.functional.gelu(x, approximate="tanh")
                acts.append(x.to(dt))
            layers.append({"name": f"blocks.{b}.{name}", "weight": w, "activations": acts})
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = search.search_layers(layers, rank=rank, world=world)
    torch.cuda.synchronize()
    dt_s = time.perf_counter() - t0
    # the same table with each layer's calibration set row-stacked into one matrix (large GEMMs, few launches)
    t0 = time.perf_counter()
    res_b = search.search_layers(layers, rank=rank, world=world, layer_fn=search.search_layer_batched)
    torch.cuda.synchronize()
    dt_b = time.perf_counter() - t0
    agree = sum(a["weight_format"] == b["weight_format"] and a["activation_format"] == b["activation_format"] for a, b in zip(res, res_b))
    rel = max(abs(a["loss"] - b["loss"]) / a["loss"] for a, b in zip(res, res_b))
    # the reference's loop (search_fp4_format.py:798-816) with the reference's quantizer path (its glue, restated in
    # tests/ref_glue.py, around its unmodified extension): every pair re-quantizes x and recomputes y_fp
    dt_r = None
    if os.environ.get("SEARCH_REF", "1") == "1" and rank == 0:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
        import ref_glue
        from fpqvar_b200 import quant_utils as Q
        ext = ref_glue.load_ref_ext()
        if ext is not None:
            grids = {"e1m2": Q.fp4_e1m2_grid, "e2m1": Q.fp4_e2m1_grid, "e3m0": Q.fp4_e3m0_grid}
            t0 = time.perf_counter()
            for ly in layers:
                w = ly["weight"]
                for wf in search.FP4_FORMATS:
                    wq = ref_glue.sym_group_cuda(ext.quant, w, grids[wf])
                    for af in search.FP4_FORMATS:
                        loss = 0.0
                        for x in ly["activations"]:
                            xq = ref_glue.sym_group_cuda(ext.quant, x, grids[af])
                            loss += search.compute_quant_error(torch.matmul(x, w.T), torch.matmul(xq, wq.T))
                        loss = float(loss / len(ly["activations"]))
            torch.cuda.synchronize()
            dt_r = time.perf_counter() - t0
    # tensor-level scores of the same activations (one kernel per tensor scores all three candidates)
    t1 = time.perf_counter()
    for ly in layers:
        for x in ly["activations"]:
            if x.numel() % 128 == 0:
                search.score_tensor_formats(x, search.FP4_FORMATS)
    torch.cuda.synchronize()
    dt_t = time.perf_counter() - t1
    if world > 1:
        import torch.distributed as dist
        objs = [None] * world
        dist.all_gather_object(objs, res)
        assert all(o == objs[0] for o in objs), "ranks disagree on the optimum"
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps({"layers": len(layers), "activations_per_layer": n_act, "world": world, "output_level_search_s": round(dt_s, 3),
                          "output_level_search_batched_s": round(dt_b, 3), "batched_same_optimum": f"{agree}/{len(res)}",
                          "batched_max_rel_loss_diff": rel, "reference_loop_reference_quantizers_s": None if dt_r is None else round(dt_r, 3),
                          "tensor_level_scoring_s": round(dt_t, 3), "first": res[:4]}))


if __name__ == "__main__":
    main()
