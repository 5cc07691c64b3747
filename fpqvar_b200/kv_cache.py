"""KV-cache fake quantization without re-quantizing the history.

The reference, with `--quant_kv`, re-quantizes the WHOLE cache at every scale before appending the new keys / values
(models_fp_quant_transform_rotate/basic_var.py:188-203):

    if quant_KV:  self.cached_k = fp6_quant_e2m3_per_token_cuda(self.cached_k, 6)     # or fp_quant_e2_per_group_cuda(.., 4)
    k = self.cached_k = torch.cat((self.cached_k, k), dim=dim_cat)

so a row appended at scale t is quantized at scales t+1, t+2, ...: sum_t cur_L(t) = 1030 token rows per block for the 256x256
schedule, where 424 suffice if each row is quantized once.  That is bit-identical to the reference because the fp16 fake
quantizer is IDEMPOTENT: a second pass re-derives the same scale from the quantized absmax element and maps every value to
itself.  Checked exhaustively against the oracle for every fp16 (absmax, x) pair (tools/idempotence_check.py); the only
exceptions are rows whose absmax is below 2^-17 (scales in the deep fp16-subnormal range) or equal to 65504 (the quantized
absmax overflows to inf).  `IncrementalKVQuant` watches for those rows on the device and, if one ever appears, `exact` turns
False: the caller can then redo the pass with `incremental=False`, which is the reference's schedule verbatim.

Layout: the appended dimension must be the outermost after the batch ("BLHc", what the reference's flash-attention path
uses, dim_cat=1): then per-token rows of 64 (kv_bit=6) and groups of 128 along the flattened head dims (kv_bit=4) never
straddle an append boundary.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import quant_utils

_TINY = 2.0 ** -17          # below this absmax a row is not idempotent under e2m3 (5.6e-6) / e2m1 (5.4e-7)
_HUGE = 65504.0


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
        self._suspect: Optional[torch.Tensor] = None      # device flag: a row outside the idempotent range was quantized

    def reset(self):
        self.cur = self.done = 0
        if self._suspect is not None:
            self._suspect.zero_()

    @property
    def exact(self) -> bool:
        """False if a row outside the proven-idempotent range went through the incremental schedule (reads a device flag:
        call it once, after the pass)."""
        return self._suspect is None or not bool(self._suspect.item())

    def _watch(self, t: torch.Tensor):
        if self.kv_bit == 6:
            amax = t.abs().amax(dim=-1)
        else:
            amax = t.reshape(-1, 128).abs().amax(dim=-1)
        bad = ((amax < _TINY) & (amax > 0)) | (amax >= _HUGE) | torch.isnan(amax)
        self._suspect |= bad.any()

    def append(self, k: torch.Tensor, v: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        if k.shape != v.shape or k.dim() != 4 or k.dtype != torch.float16:
            raise ValueError("IncrementalKVQuant.append: k and v must be fp16 [B, l, H, head_dim] of the same shape")
        B, l, H, hd = k.shape
        if self.kv_bit == 4 and (H * hd) % 128:
            raise ValueError("kv_bit=4 quantizes groups of 128 along the flattened [H, head_dim]: H * head_dim must be a multiple of 128")
        if self.k is None or self.k.shape[0] != B or self.k.shape[2:] != (H, hd) or self.k.device != k.device:
            self.k = torch.empty(B, self.max_len, H, hd, dtype=torch.float16, device=k.device)
            self.v = torch.empty_like(self.k)
            self._suspect = torch.zeros((), dtype=torch.bool, device=k.device)
            self.cur = self.done = 0
        if self.cur + l > self.max_len:
            raise ValueError(f"cache overflow: {self.cur} + {l} > {self.max_len}")
        if self.cur:                                        # basic_var.py:189-203: quantize what is cached, then append
            lo = self.done if self.incremental else 0
            for buf in (self.k, self.v):
                part = buf[:, lo:self.cur]
                if self.incremental:
                    self._watch(part)
                part.copy_(_quant(part, self.kv_bit))
            self.done = self.cur
        self.k[:, self.cur:self.cur + l] = k
        self.v[:, self.cur:self.cur + l] = v
        self.cur += l
        return self.k[:, :self.cur], self.v[:, :self.cur]
