"""Shapes of the fake-quant hot path inside one VAR generation pass (host logic, no CUDA).

One autoregressive pass of the reference (models_fp_quant_transform_rotate/var.py:135-217)
runs 10 next-scale stages; stage ``si`` feeds ``2*B*pn[si]^2`` token rows (CFG doubles the
batch, var.py:166,214) through ``depth`` transformer blocks, and every block calls an
activation quantizer four times (QuantizedLinear.forward, quant_utils.py:764-769,991-996):

    site      input                                  reference op
    mat_qkv   adaLN-modulated LN output  [rows, C]   .mul(s_qkv) @ Q  -> fp_quant_e2_per_group_cuda   (basic_var.py:263)
    proj      attention output           [rows, C]   fp_quant_e2_per_group_cuda
    fc1       adaLN-modulated LN output  [rows, C]   .mul(s_fc1) @ Q  -> fp_quant_e2_per_group_cuda   (basic_var.py:266)
    fc2       GELU(tanh) output          [rows, 4C]  fc2 act type (fp_e1m2_neg_e2m1_pos for the README command)

``C = 64*depth`` (models*/__init__.py:19-20).  Without --rotate/--transform (models_fp_quant) the
mat_qkv / fc1 inputs are quantized as they are, in fp32.

The list produced here is what bench.py replays, what the multi-GPU sharding splits, and what
DESIGN.md's algorithmic-byte figures are computed from.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

PATCH_NUMS_256 = (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)       # evaluate_fp_quant_transform_rotate.py:63
PATCH_NUMS_512 = (1, 2, 3, 4, 6, 9, 13, 18, 24, 32)      # evaluate_fp_quant_transform_rotate_512x512.py:62


@dataclass(frozen=True)
class Call:
    """One activation-quantizer call.  ``op``: "group" (symmetric per-group), "signsplit",
    "rotate_quant" (GALT multiply + block Hadamard + per-group quant), "mod_rotate_quant" (the same with the adaLN
    modulate ``x * (scale + 1) + shift`` in front of it, basic_var.py:263,266: what INTEGRATION.md level (c) runs at
    mat_qkv / fc1; ``rows_per_batch`` token rows share one [C] row of scale / shift)."""
    stage: int
    block: int
    site: str
    op: str
    fmt: str
    rows: int
    cols: int
    in_dtype: str        # "f32" | "f16"
    out_dtype: str
    rows_per_batch: int = 0

    @property
    def elems(self) -> int:
        return self.rows * self.cols

    @property
    def in_bytes(self) -> int:
        return self.elems * (4 if self.in_dtype == "f32" else 2)

    @property
    def out_bytes(self) -> int:
        return self.elems * (4 if self.out_dtype == "f32" else 2)

    @property
    def aux_bytes(self) -> int:
        """The fp32 adaLN scale and shift rows a "mod_rotate_quant" call reads: 2 x [rows / rows_per_batch, cols]."""
        return 2 * (self.rows // self.rows_per_batch) * self.cols * 4 if self.op == "mod_rotate_quant" else 0

    @property
    def bytes(self) -> int:
        """Algorithmic bytes: every input element (and adaLN operand) read once, every output element written once."""
        return self.in_bytes + self.out_bytes + self.aux_bytes


@dataclass(frozen=True)
class VarHotPath:
    name: str
    depth: int
    batch: int                       # images per pass (rows are doubled by CFG)
    patch_nums: Sequence[int]
    rotate_transform: bool           # --rotate --block_rotate --transform
    act_fmt: str = "e2m1"            # fp_e2
    fc2_op: str = "signsplit"
    fc2_fmt: str = "e1m2_neg_e2m1_pos"
    modulate: bool = False           # mat_qkv / fc1: the adaLN modulate fused into the rotate kernel (needs rotate_transform)

    @property
    def width(self) -> int:
        return 64 * self.depth

    def stage_rows(self) -> List[int]:
        return [2 * self.batch * pn * pn for pn in self.patch_nums]

    def calls(self, blocks: Sequence[int] | None = None) -> List[Call]:
        c = self.width
        out: List[Call] = []
        rot_op = "mod_rotate_quant" if self.modulate else "rotate_quant"
        for si, rows in enumerate(self.stage_rows()):
            rpb = self.patch_nums[si] ** 2 if self.modulate else 0        # token rows of one image at this stage
            for b in (range(self.depth) if blocks is None else blocks):
                if self.rotate_transform:
                    out.append(Call(si, b, "mat_qkv", rot_op, self.act_fmt, rows, c, "f32", "f16", rpb))
                else:
                    out.append(Call(si, b, "mat_qkv", "group", self.act_fmt, rows, c, "f32", "f32"))
                out.append(Call(si, b, "proj", "group", self.act_fmt, rows, c, "f16", "f16"))
                if self.rotate_transform:
                    out.append(Call(si, b, "fc1", rot_op, self.act_fmt, rows, c, "f32", "f16", rpb))
                else:
                    out.append(Call(si, b, "fc1", "group", self.act_fmt, rows, c, "f32", "f32"))
                out.append(Call(si, b, "fc2", self.fc2_op, self.fc2_fmt, rows, 4 * c, "f16", "f16"))
        return out

    def bytes_per_pass(self) -> int:
        return sum(k.bytes for k in self.calls())

    def elems_per_pass(self) -> int:
        return sum(k.elems for k in self.calls())


# BASELINE.json configs[1..3]
WORKLOADS = {
    # configs[1]: VAR-d16 256x256 W4A4 fp_e2 per-group, batch 64 (models_fp_quant: no rotation)
    "var_d16_w4a4": VarHotPath("var_d16_w4a4", 16, 64, PATCH_NUMS_256, False, "e2m1", "group", "e2m1"),
    # configs[2]: VAR-d30 256x256 W4A4 fp_e2 + fc2 fp_e1m2_neg_e2m1_pos, block rotate + GALT (README.md:33), B=50 per GPU;
    # mat_qkv / fc1 through the adaLN-fused kernel, as INTEGRATION.md level (c) calls it
    "var_d30_w4a4_rot": VarHotPath("var_d30_w4a4_rot", 30, 50, PATCH_NUMS_256, True, "e2m1", "signsplit", "e1m2_neg_e2m1_pos", True),
    # configs[3]: VAR-d36 512x512 W6A6 FP6 per-group, rotate + transform, B=10 per call
    "var_d36_w6a6_rot": VarHotPath("var_d36_w6a6_rot", 36, 10, PATCH_NUMS_512, True, "e2m3", "signsplit", "int_neg_e2m3_pos", True),
    # the same two with the modulate left to the caller (three ATen kernels in front of `.mul(s) @ Q`): round-1's step
    "var_d30_w4a4_rot_nomod": VarHotPath("var_d30_w4a4_rot_nomod", 30, 50, PATCH_NUMS_256, True, "e2m1", "signsplit", "e1m2_neg_e2m1_pos"),
    "var_d36_w6a6_rot_nomod": VarHotPath("var_d36_w6a6_rot_nomod", 36, 10, PATCH_NUMS_512, True, "e2m3", "signsplit", "int_neg_e2m3_pos"),
}


def shard_units(n_units: int, rank: int, world: int) -> range:
    """Round-robin ownership of independent units (classes / image batches / search candidates):
    unit i belongs to rank i % world."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return range(rank, n_units, world)
