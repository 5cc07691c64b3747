"""KV-cache fake quantization without re-quantizing the history.

The reference, with `--quant_kv`, re-quantizes the WHOLE cache at every scale before appending the new keys / values
(models_fp_quant_transform_rotate/basic_var.py:188-203):

    if quant_KV:  self.cached_k = fp6_quant_e2m3_per_token_cuda(self.cached_k, 6)     # or fp_quant_e2_per_group_cuda(.., 4)
    k = self.cached_k = torch.cat((self.cached_k, k), dim=dim_cat)

so a row appended at scale t is quantized at scales t+1, t+2, ...: sum_t cur_L(t) = 1030 token rows per block for the 256x256
schedule, where 424 suffice if each row is quantized once.  That is bit-identical to the reference because the fp16 fake
quantizer is IDEMPOTENT: a second pass re-derives the same scale from the quantized absmax element and maps every value to
itself.  Checked exhaustively against the oracle for every fp16 (absmax, x) pair (tests/idempotence_exhaustive.py); the only
exceptions are rows whose absmax is below 2^-17 (scales in the deep fp16-subnormal range) or equal to 65504 (the quantized
absmax overflows to inf).  Such rows are recognisable afterwards (a quantized row with 0 < absmax < 2^-16, or a non-finite
value): `IncrementalKVQuant.exact` scans the finished cache once and reports them, and the caller can redo the pass with
`incremental=False`, which is the reference's schedule verbatim.  Nothing is checked inside `append` (a per-append check
costs more launches than the quantizer itself on the launch-bound early scales).

Layout: the appended dimension must be the outermost after the batch ("BLHc", what the reference's flash-attention path
uses, dim_cat=1): then per-token rows of 64 (kv_bit=6) and groups of 128 along the flattened head dims (kv_bit=4) never
straddle an append boundary.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops, quant_utils

_TINY = 2.0 ** -16          # rows that drift have absmax <= 5.6e-6 (e2m3) / 5.4e-7 (e2m1) before AND after quantization: < 2^-17


def _quant(t: torch.Tensor, kv_bit: int) -> torch.Tensor:
    if kv_bit == 6:
        return quant_utils.fp6_quant_e2m3_per_token_cuda(t, kv_bit)          # basic_var.py:193-194
    if kv_bit == 4:
        return quant_utils.fp_quant_e2_per_group_cuda(t, kv_bit)             # basic_var.py:196-197
    raise NotImplementedError                                               # basic_var.py:199


class IncrementalKVQuant:
    """One attention layer's K and V cache, [B, L_max, H, head_dim] fp16, filled scale by scale.

        cache = IncrementalKVQuant(kv_bit=6, max_len=680)
        k_all, v_all = cache.append(k, v)        # k, v: [B, l, H, head_dim]; returns views of the first cur_L tokens

    `append` returns exactly what the reference's `self.cached_k` / `self.cached_v` hold after the same call: every
    token of earlier scales fake-quantized, the tokens of this scale as they came.  incremental=False re-quantizes the
    whole history at every call (the reference's schedule)."""

    def __init__(self, kv_bit: int, max_len: int, incremental: bool = True):
        if kv_bit not in (4, 6):
            raise NotImplementedError
        self.kv_bit, self.max_len, self.incremental = kv_bit, max_len, incremental
        self.k: Optional[torch.Tensor] = None
        self.v: Optional[torch.Tensor] = None
        self.cur = 0            # tokens in the cache
        self.done = 0           # tokens already quantized

    def reset(self):
        self.cur = self.done = 0

    @property
    def exact(self) -> bool:
        """True if every row quantized so far lies in the range where one pass equals the reference's repeated passes.
        Scans the quantized part of the cache (a few reductions and one host read): call it once, after the pass."""
        if not self.incremental or self.k is None or self.done == 0:
            return True
        bad = False
        for buf in (self.k, self.v):
            part = buf[:, :self.done]
            amax = (part.abs().amax(dim=-1) if self.kv_bit == 6 else part.reshape(-1, 128).abs().amax(dim=-1)).float()
            bad = bad | (((amax < _TINY) & (amax > 0)) | ~torch.isfinite(amax)).any()
        return not bool(bad)

    def _quant_inplace(self, buf: torch.Tensor, lo: int, hi: int) -> None:
        """basic_var.py:193-197 on buf[:, lo:hi], where the rows lie: one launch, one read and one write of the slice (the
        slice is B pieces of a larger tensor: fpq_fake_quant_segments walks them with a pitch).  kv_bit=6: per-token rows of
        head_dim; kv_bit=4: groups of 128 along the flattened [H, head_dim]."""
        hd = buf.shape[-1]
        if self.kv_bit == 6 and hd == 64:
            ops.fake_quant_segments_(buf, lo, hi, "e2m3", 64)
        elif self.kv_bit == 4:
            ops.fake_quant_segments_(buf, lo, hi, "e2m1", 128)
        else:                                   # other head sizes: the general row kernel on a gathered copy
            part = buf[:, lo:hi]
            part.copy_(_quant(part, self.kv_bit))

    def append(self, k: torch.Tensor, v: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        if k.shape != v.shape or k.dim() != 4 or k.dtype != torch.float16:
            raise ValueError("IncrementalKVQuant.append: k and v must be fp16 [B, l, H, head_dim] of the same shape")
        B, l, H, hd = k.shape
        if self.kv_bit == 4 and (H * hd) % 128:
            raise ValueError("kv_bit=4 quantizes groups of 128 along the flattened [H, head_dim]: H * head_dim must be a multiple of 128")
        if self.k is None or self.k.shape[0] != B or self.k.shape[2:] != (H, hd) or self.k.device != k.device:
            self.k = torch.empty(B, self.max_len, H, hd, dtype=torch.float16, device=k.device)
            self.v = torch.empty_like(self.k)
            self.cur = self.done = 0
        if self.cur + l > self.max_len:
            raise ValueError(f"cache overflow: {self.cur} + {l} > {self.max_len}")
        if self.cur:                                        # basic_var.py:189-203: quantize what is cached, then append
            lo = self.done if self.incremental else 0
            for buf in (self.k, self.v):
                self._quant_inplace(buf, lo, self.cur)
            self.done = self.cur
        self.k[:, self.cur:self.cur + l] = k
        self.v[:, self.cur:self.cur + l] = v
        self.cur += l
        return self.k[:, :self.cur], self.v[:, :self.cur]
