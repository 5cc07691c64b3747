#!/usr/bin/env python
"""Per-token / per-channel (row) kernels at the README FP6 shapes (development aid): rows of C and 4C of VAR-d30 / d36,
symmetric e2m3 and sign-split int_neg_e2m3_pos, fp16 -> fp16, for every values-per-thread choice (tunable row_v).
Rotating buffers (>> L2), CUDA events, algorithmic GB/s (4 B/element)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fpqvar_b200 import _lib as L  # noqa: E402

if os.environ.get("FPQ_LIB_PATH"):                     # an alternative build (fpqvar_b200/csrc: make VARIANT=... EXTRA=...)
    L.LIB_PATH = os.environ["FPQ_LIB_PATH"]
lib = L.lib()
dev = torch.device("cuda")
st = torch.cuda.current_stream().cuda_stream
NBUF, ITERS = 6, int(os.environ.get("KB_ITERS", "30"))
TOKENS = int(os.environ.get("ROWS", "25600"))


def timeit(fn):
    for i in range(NBUF):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(ITERS):
        fn(i % NBUF)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3 / ITERS


def main():
    print(f"{'row_len':>8} {'rows':>7} {'kernel':<28} " + " ".join(f"{'V=' + v:>10}" for v in ("auto", "1", "2", "4")) + "   (GB/s)")
    for row_len in (1920, 2304, 7680, 9216):
        rows = TOKENS if row_len <= 2304 else TOKENS * 1920 // row_len * 4 // 4
        rows = min(rows, (3 << 29) // (row_len * 2 * NBUF))                # <= 1.5 GiB of inputs
        x = [torch.nn.functional.gelu(torch.randn(rows, row_len, device=dev)).half() for _ in range(NBUF)]
        o = [torch.empty_like(t) for t in x]
        for name, fn in (("sym e2m3", lambda i: lib.fpq_fake_quant(x[i].data_ptr(), o[i].data_ptr(), rows, row_len, 1, 1, 3, 0, 0, st)),
                         ("split int_neg_e2m3_pos", lambda i: lib.fpq_fake_quant_signsplit(x[i].data_ptr(), o[i].data_ptr(), rows, row_len, 1, 1, 1, 0, 0, None, st))):
            cells = []
            for v in ("", "1", "2", "4"):
                lib.fpq_set_tunable(b"row_v", int(v) if v else 0)
                if v and (row_len // 8 + int(v) - 1) // int(v) > 1024:
                    cells.append(f"{'-':>10}")
                    continue
                t = timeit(fn)
                cells.append(f"{rows * row_len * 4 / t / 1e9:10.0f}")
            lib.fpq_set_tunable(b"row_v", 0)
            print(f"{row_len:>8} {rows:>7} {name:<28} " + " ".join(cells), flush=True)
        del x, o
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
