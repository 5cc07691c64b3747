"""Host-side model of the streaming rotate kernel's shared-memory choreography (fpqvar_b200/csrc/fpq_rotate.cu,
rotate_tma_body): the launch plan (rot_plan), which lane touches which 16-byte unit in pass 1 / pass 2, the XOR swizzle,
and the butterfly order.  Used by tests/test_host_logic.py to check, without a GPU, that

  * every chunk of a tile is read by exactly one pass-1 lane set and one pass-2 lane set, and every unit once;
  * every LDS.128 / STS.128 phase (8 consecutive lanes) touches 8 distinct 16-byte bank slots (conflict-free);
  * the two passes together compute FWHT_128 in the butterfly order 0,1,5,6,2,3,4.

Test infrastructure only: nothing in fpqvar_b200/ imports it."""
from __future__ import annotations

import numpy as np

ROT_STAGES = 3
ROT_MAX_WARPS = 9


def rot_plan(cpr: int):
    """Mirror of rot_plan() in fpq_rotate.cu.  Returns dict or None."""
    best, best_util = None, 0.0
    cw = 4
    while cw >= 1:
        wcols = (cpr + cw - 1) // cw
        if wcols <= ROT_MAX_WARPS:
            wr = ROT_MAX_WARPS // wcols
            if wcols * wr > 8 and wr > 1:
                wr = 8 // wcols if 8 // wcols > 0 else 1
            while wr >= 1:
                rs = wr * 2 * (4 // cw)
                stage = rs * cpr * 512
                if stage * ROT_STAGES + 2 * ROT_STAGES * 8 > 110 * 1024:
                    wr -= 1
                    continue
                util = cpr / (wcols * cw)
                if best is None or util > best_util + 1e-9:
                    best = dict(cpr=cpr, cw=cw, wcols=wcols, n_warps=wcols * wr, rs=rs, stage_bytes=stage)
                    best_util = util
                break
        cw >>= 1
    return best


def lane_maps(plan, warp: int, lane: int):
    """Per-lane constants exactly as the kernel derives them."""
    cw, wcols, cpr = plan["cw"], plan["wcols"], plan["cpr"]
    wc, wr = warp % wcols, warp // wcols
    rps = 4 // cw
    g4, l8 = lane >> 3, lane & 7
    col1 = wc * cw + g4 % cw
    trow1 = wr * 2 * rps + g4 // cw
    s1 = g4 & 1
    g8, lq = lane >> 2, lane & 3
    col2 = wc * cw + (g8 & 3) % cw
    trow2 = wr * 2 * rps + (g8 >> 2) * rps + (g8 & 3) // cw
    key2 = 2 * lq + (g8 & 1)
    return dict(rps=rps, g4=g4, l8=l8, col1=col1, trow1=trow1, s1=s1, g8=g8, lq=lq, col2=col2, trow2=trow2, key2=key2)


def pass1_accesses(plan, warp, lane, sub):
    """[(load byte offset in stage, store byte offset in stage)] for a = 0..3, or None if the lane is idle."""
    m = lane_maps(plan, warp, lane)
    if m["col1"] >= plan["cpr"]:
        return None
    trow = m["trow1"] + sub * m["rps"]
    if trow >= plan["rs"]:
        return None
    cb = (trow * plan["cpr"] + m["col1"]) * 512
    out = []
    for a in range(4):
        ld = cb + a * 128 + m["l8"] * 16
        slot = (m["l8"] ^ m["s1"]) ^ (2 * a)
        st = cb + a * 128 + slot * 16
        out.append((ld, st))
    return out


def pass2_accesses(plan, warp, lane):
    m = lane_maps(plan, warp, lane)
    if m["col2"] >= plan["cpr"] or m["trow2"] >= plan["rs"]:
        return None
    cb = (m["trow2"] * plan["cpr"] + m["col2"]) * 512 + m["lq"] * 128
    return [cb + ((i ^ m["key2"]) << 4) for i in range(8)]


def conflict_free(addrs):
    """addrs: byte addresses of the 8 lanes of one LDS.128/STS.128 phase (None = inactive lane)."""
    slots = [(a // 16) % 8 for a in addrs if a is not None]
    return len(slots) == len(set(slots))


def simulate_tile(plan, tile: np.ndarray, mult: np.ndarray) -> np.ndarray:
    """tile: [rs, cpr*128] fp64 (or fp32).  Runs pass 1 and pass 2 exactly as the kernel orders its butterflies and moves
    its data through the (swizzled) stage; returns the rotated tile [rs, cpr*128]."""
    rs, cpr = plan["rs"], plan["cpr"]
    stage = np.zeros(rs * cpr * 128, dtype=tile.dtype)
    stage[:] = tile.reshape(-1)
    out = np.full_like(tile, np.nan)
    for warp in range(plan["n_warps"]):
        # pass 1: every lane of the warp has its values in registers before any lane stores (same instruction stream)
        for sub in range(2):
            pending = []
            for lane in range(32):
                acc = pass1_accesses(plan, warp, lane, sub)
                if acc is None:
                    continue
                m = lane_maps(plan, warp, lane)
                v = np.zeros(16, dtype=tile.dtype)
                for a, (ld, _) in enumerate(acc):
                    v[4 * a:4 * a + 4] = stage[ld // 4: ld // 4 + 4]
                    e0 = 32 * a + 4 * m["l8"]
                    v[4 * a:4 * a + 4] *= mult[m["col1"] * 128 + e0: m["col1"] * 128 + e0 + 4]
                # bits 0, 1 inside a unit; bits 5, 6 = a
                for a in range(4):
                    u = v[4 * a:4 * a + 4].copy()
                    u = np.array([u[0] + u[1], u[0] - u[1], u[2] + u[3], u[2] - u[3]])
                    u = np.array([u[0] + u[2], u[1] + u[3], u[0] - u[2], u[1] - u[3]])
                    v[4 * a:4 * a + 4] = u
                w = v.reshape(4, 4)
                w = np.stack([w[0] + w[1], w[0] - w[1], w[2] + w[3], w[2] - w[3]])
                w = np.stack([w[0] + w[2], w[1] + w[3], w[0] - w[2], w[1] - w[3]])
                pending.append((acc, w))
            for acc, w in pending:
                for a, (_, st) in enumerate(acc):
                    stage[st // 4: st // 4 + 4] = w[a]
        # pass 2
        for lane in range(32):
            acc = pass2_accesses(plan, warp, lane)
            if acc is None:
                continue
            m = lane_maps(plan, warp, lane)
            q = np.stack([stage[a // 4: a // 4 + 4] for a in acc])       # [8 units i, 4]
            for h in (1, 2, 4):                                           # bits 2, 3, 4
                n = q.copy()
                for i in range(8):
                    if i & h == 0:
                        n[i] = q[i] + q[i + h]
                        n[i + h] = q[i] - q[i + h]
                q = n
            c0 = m["col2"] * 128 + m["lq"] * 32
            out[m["trow2"], c0:c0 + 32] = q.reshape(-1)
    return out
