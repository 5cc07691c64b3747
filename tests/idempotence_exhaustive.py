#!/usr/bin/env python
"""Is the fp16 fake quantizer idempotent?  (Decides whether the KV cache can be quantized incrementally)

Exhaustive over the oracle: for EVERY finite fp16 absmax a and EVERY fp16 x with |x| <= a (the quantizer is elementwise once
the scale is fixed, and the second-pass scale depends only on the image of the absmax element), quantize, re-derive the scale
from the quantized row, quantize again, compare bits.  ~4 min on one core.  Result (committed in DESIGN.md section 8):
  e2m3 per_token: idempotent except absmax in {3.40e-6, 3.81e-6, 3.87e-6, 4.29e-6, 4.71e-6, 5.60e-6} (scales < 2^-20) and 65504
  e2m1 per_group: idempotent except absmax 5.36e-7 and 65504 (half(6 * s) overflows)
"""
import numpy as np, sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))  # TEST INFRASTRUCTURE: runs the oracle, not the product
from oracle import oracle as O
allpos=np.arange(0x0000,0x7C00,dtype=np.uint16).view(np.float16)   # +0 .. max finite
def Q(x, s, grid, vmaxf):
    with np.errstate(all='ignore'):
        v=(x.astype(np.float32)/s.astype(np.float32)).astype(np.float16)
        q=O.scan_quant(v.astype(np.float32), grid)
        return (q*s.astype(np.float32)).astype(np.float16)
for fmt in ("e2m3","e2m1"):
    grid=O.GRIDS[fmt]; vmax=np.float32(O.grid_absmax(fmt))
    t0=time.time(); bad_a=[]
    for ai in range(1,0x7C00):
        a=allpos[ai]
        with np.errstate(all='ignore'):
            s=(np.float32(a)/vmax).astype(np.float16)
        x=allpos[:ai+1]                      # every |x| <= a (sign symmetric)
        y=Q(x, np.float16(s), grid, vmax)
        a2=np.max(np.abs(y[np.isfinite(y)])) if np.isfinite(y).any() else np.float16(0)
        if not np.isfinite(y).all(): bad_a.append((ai,'nonfinite')); continue
        with np.errstate(all='ignore'):
            s2=(np.float32(a2)/vmax).astype(np.float16)
        y2=Q(y, np.float16(s2), grid, vmax)
        if not np.array_equal(y.view(np.uint16), y2.view(np.uint16)):
            bad_a.append((ai, float(a), float(s), float(s2), int((y.view(np.uint16)!=y2.view(np.uint16)).sum())))
    print(fmt, "absmax values whose rows are not idempotent:", len(bad_a), "time", round(time.time()-t0,1))
    print(bad_a[:10], bad_a[-5:])
