"""Host-side model of the streaming rotate kernel's shared-memory choreography (fpqvar_b200/csrc/fpq_rotate.cu,
rotate_stream_body): the launch plan (rot_plan), which lane touches which 16-byte unit of the warp's 4 KB buffer in
pass 1 / pass 2, the XOR swizzle, and the butterfly order.  Used by tests/test_host_logic.py to check, without a GPU, that

  * every unit of the warp's 2 rows x 4 chunks is touched by exactly one pass-1 lane and one pass-2 lane;
  * every LDS.128 / STS.128 phase (8 consecutive lanes) touches 8 distinct 16-byte bank slots (conflict-free);
  * pass 1 writes back strictly inside the units its own chunk's lanes loaded (in place);
  * the two passes together compute FWHT_128 in the butterfly order 0,1,5,6,2,3,4.

Test infrastructure only: nothing in fpqvar_b200/ imports it."""
from __future__ import annotations

import numpy as np

ROT_MIN_CPR, ROT_MAX_CPR = 4, 36
STAGE_BYTES = 4096                      # 2 rows x 4 chunks x 512 B per warp and buffer


def rot_plan(cpr: int, mod: bool = False):
    """Mirror of rot_plan() in fpq_rotate.cu: (warps per CTA, CTAs per SM) or None."""
    if cpr < ROT_MIN_CPR or cpr > ROT_MAX_CPR:
        return None
    warps = (cpr + 3) // 4
    return warps, max(1, min((16 if mod else 20) // warps, 228 * 1024 // (warps * (2 * 4096 + 16) + 1024)))


def pass1_accesses(lane: int, sub: int, ncols: int = 4, nr: int = 2):
    """[(load byte offset, store byte offset)] in the warp's buffer for a = 0..3, or None if the lane is idle."""
    g4, l8 = lane >> 3, lane & 7
    if g4 >= ncols or sub >= nr:
        return None
    cb = sub * 2048 + g4 * 512
    slot1 = l8 ^ (g4 & 1)
    return [(cb + a * 128 + l8 * 16, cb + a * 128 + (slot1 ^ (2 * a)) * 16) for a in range(4)]


def pass2_accesses(lane: int, ncols: int = 4, nr: int = 2):
    g8, lq = lane >> 2, lane & 3
    row2, c2 = g8 >> 2, g8 & 3
    if c2 >= ncols or row2 >= nr:
        return None
    cb = row2 * 2048 + c2 * 512 + lq * 128
    key2 = 2 * lq + (g8 & 1)
    return [cb + ((i ^ key2) << 4) for i in range(8)]


def conflict_free(addrs) -> bool:
    """addrs: byte addresses of the 8 lanes of one LDS.128 / STS.128 phase (None = inactive lane)."""
    slots = [(a // 16) % 8 for a in addrs if a is not None]
    return len(slots) == len(set(slots))


def simulate_step(tile: np.ndarray, mult: np.ndarray, ncols: int = 4) -> np.ndarray:
    """tile: [nr <= 2, ncols * 128]; mult: [ncols * 128].  Runs pass 1 and pass 2 exactly as the kernel orders its
    butterflies and moves its data through the (swizzled) buffer; returns the rotated rows."""
    nr = tile.shape[0]
    stage = np.zeros(STAGE_BYTES // 4, dtype=tile.dtype)
    for j in range(nr):                                     # one bulk copy per row: 2048-byte pitch
        stage[j * 512: j * 512 + ncols * 128] = tile[j]
    out = np.full_like(tile, np.nan)
    # pass 1: every lane of a chunk has its values in registers before any of them stores (__syncwarp(mask))
    for sub in range(2):
        pending = []
        for lane in range(32):
            acc = pass1_accesses(lane, sub, ncols, nr)
            if acc is None:
                continue
            g4, l8 = lane >> 3, lane & 7
            v = np.zeros((4, 4), dtype=tile.dtype)
            for a, (ld, _) in enumerate(acc):
                e0 = 32 * a + 4 * l8
                v[a] = stage[ld // 4: ld // 4 + 4] * mult[g4 * 128 + e0: g4 * 128 + e0 + 4]
            for a in range(4):                              # bits 0, 1 inside a unit
                u = v[a]
                u = np.array([u[0] + u[1], u[0] - u[1], u[2] + u[3], u[2] - u[3]])
                v[a] = np.array([u[0] + u[2], u[1] + u[3], u[0] - u[2], u[1] - u[3]])
            w = np.stack([v[0] + v[1], v[0] - v[1], v[2] + v[3], v[2] - v[3]])          # bit 5 = a bit 0
            w = np.stack([w[0] + w[2], w[1] + w[3], w[0] - w[2], w[1] - w[3]])          # bit 6 = a bit 1
            pending.append((acc, w))
        for acc, w in pending:
            for a, (_, st) in enumerate(acc):
                stage[st // 4: st // 4 + 4] = w[a]
    # pass 2
    for lane in range(32):
        acc = pass2_accesses(lane, ncols, nr)
        if acc is None:
            continue
        g8, lq = lane >> 2, lane & 3
        q = np.stack([stage[a // 4: a // 4 + 4] for a in acc])                           # [8 units i, 4]
        for h in (1, 2, 4):                                                               # bits 2, 3, 4
            n = q.copy()
            for i in range(8):
                if i & h == 0:
                    n[i] = q[i] + q[i + h]
                    n[i + h] = q[i] - q[i + h]
            q = n
        c0 = (g8 & 3) * 128 + lq * 32
        out[g8 >> 2, c0:c0 + 32] = q.reshape(-1)
    return out
