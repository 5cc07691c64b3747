// Packed fast path for fp16 tensors under the kernel tie rule (the activation flow of the
// rotated models: fp16 in, fp16 out).
//
// The generic path (fpq_common.cuh) spends ~20-38 instructions per element and is issue-bound on
// B200 (ncu: 85 % issue-slot utilisation at 36 % DRAM, profiles/r1a_generic_kernels_ncu.txt).  At 4 bytes per
// element the HBM roofline leaves ~20 issue slots per element, so this path is written for
// instruction count (~8 per element):
//   * two elements per instruction wherever the ISA allows it: FMUL2 / FFMA2 (packed fp32,
//     sm_100), F2FP.PACK_AB, VIMNMX3 on 16x2 lanes;
//   * grid rounding without a division, a directed-rounding add or integer masking per element:
//       w = v + delta          v = half(x * RN(1/s)) == half(x/s) (proof: fpq_common.cuh), widened by the
//                              mixed-precision add FHADD (add.f32.f16), which adds delta for free.
//                              delta = +2^-17 moves exact midpoints to the side the kernel rule sends
//                              them to (+inf) and is smaller than the distance from any other fp16
//                              value to a midpoint, so nothing else changes side; v has 11
//                              significant bits, so v + delta is exact in fp32.
//       p = 2^max(exponent(w), EMIN)
//       y = RN(p * 1.5*2^(23-M) + w)      the round-to-nearest-even of the FFMA does the grid rounding
//                                         (w is never a tie any more): ulp(y) = 2^(E-M)
//       q = y - p * 1.5*2^(23-M)          exact
//     which handles the subnormal region of the target format (exponent clamp) and every binade
//     with the same instructions;
//   * per-group scalars without the guarded library sequences: s = half(a * RN(1/VMAX)) (equal to
//     half(a / VMAX) for every fp16 a) and r = RN(1/s) by MUFU.RCP + one Newton step (s is a normal
//     fp16 number here).
// fpq_selftest_f16_flow checks these element and scalar functions against the literal reference
// sequence for EVERY fp16 input that can occur (tests/test_gpu_parity.py::test_f16_flow_exhaustive).
#pragma once
#include "fpq_common.cuh"

namespace fpq {

__device__ __forceinline__ float fhadd(uint16_t h, float c) {           // fp32 = fp16 + fp32, one FHADD
    float d;
    asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(d) : "h"(h), "f"(c));
    return d;
}
struct F2 { float lo, hi; };
__device__ __forceinline__ uint64_t pk(float lo, float hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ F2 unpk(uint64_t v) {
    F2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.lo), "=f"(r.hi) : "l"(v));
    return r;
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint32_t pack_h2_u64(uint64_t v) {
    const F2 f = unpk(v);
    return pack_h2(f.lo, f.hi);
}
__device__ __forceinline__ uint64_t widen_h2(uint32_t h2) {              // two halves -> two floats
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h2));
    return pk(f.x, f.y);
}

// The tie shift delta = 2^-17 of the kernel rule (header comment), kept in ONE vector register:
// FHADD takes no immediate and no uniform register, and ptxas would otherwise copy the constant
// from a uniform register in front of every FHADD.  `zero_per_thread` is a per-thread value that
// is 0 for every grid this library launches but that ptxas cannot prove uniform (callers pass the
// high bits of their 64-bit global warp index).
__device__ __forceinline__ float tie_delta_kernel(uint32_t zero_per_thread) {
    return __uint_as_float(0x37000000u + zero_per_thread);
}

template <class HG> struct Magic {
    static constexpr float EM = HG::EMIN >= 0 ? float(1u << (HG::EMIN >= 0 ? HG::EMIN : 0)) : 1.0f / float(1u << (HG::EMIN < 0 ? -HG::EMIN : 0));
    static constexpr float SC = 1.5f * float(1u << (23 - HG::M));
    static constexpr float INV_VMAX = 1.0f / HG::VMAX;         // RN(1/VMAX), evaluated by the compiler in fp32
};

// q = R_K(v) for the two halves of v2 (finite, |v| within the format's range), as packed fp32.
// `em`/`sc` may differ per element (sign-split formats).
__device__ __forceinline__ uint64_t round_pair_magic(uint32_t v2, float delta, float em0, float em1, float sc0, float sc1) {
    const float w0 = fhadd(uint16_t(v2 & 0xffffu), delta);
    const float w1 = fhadd(uint16_t(v2 >> 16), delta);
    const float p0 = __uint_as_float(__float_as_uint(fmaxf(fabsf(w0), em0)) & 0x7F800000u);
    const float p1 = __uint_as_float(__float_as_uint(fmaxf(fabsf(w1), em1)) & 0x7F800000u);
    const uint64_t p = pk(p0, p1), w = pk(w0, w1);
    const uint64_t y = ffma2(p, pk(sc0, sc1), w);
    return ffma2(p, pk(-sc0, -sc1), y);
}

// One packed pair of the symmetric flow: two fp16 inputs (already widened to packed fp32) -> two fp16
// outputs q*s.
template <class HG>
__device__ __forceinline__ uint32_t sym_pair_h16_w(uint64_t xf2, uint64_t r2, uint64_t s2, float delta) {
    const uint32_t v2 = pack_h2_u64(fmul2(xf2, r2));                              // half(x/s)
    const uint64_t q = round_pair_magic(v2, delta, Magic<HG>::EM, Magic<HG>::EM, Magic<HG>::SC, Magic<HG>::SC);
    return pack_h2_u64(fmul2(q, s2));                                             // half(q*s)
}
template <class HG>
__device__ __forceinline__ uint32_t sym_pair_h16(uint32_t x2, uint64_t r2, uint64_t s2, float delta) {
    const uint32_t v2 = pack_h2_u64(fmul2(widen_h2(x2), r2));                     // half(x/s)
    const uint64_t q = round_pair_magic(v2, delta, Magic<HG>::EM, Magic<HG>::EM, Magic<HG>::SC, Magic<HG>::SC);
    return pack_h2_u64(fmul2(q, s2));                                             // half(q*s)
}

// s = half(a / VMAX) for an fp16 absmax a (quant_utils.py:320).  a * RN(1/VMAX) is within 2^-23
// (relative) of a / VMAX, while a / VMAX (an 11-bit integer over 3, 7 or 15, or exact) stays
// >= 2^-15 (relative) away from every NORMAL fp16 rounding boundary: whenever the result is a normal
// fp16 number the two round to the same value.  (Among subnormal results a / VMAX can be an exact
// tie, e.g. 14*2^-24 / 28; those scales are "irregular" and recomputed with the true division.)
template <class HG> __device__ __forceinline__ __half scale_from_absmax_h16(float a) {
    return __float2half_rn(a * Magic<HG>::INV_VMAX);
}
// RN(1/s) for a NORMAL fp16 value s: the body of __frcp_rn without its range guard.
__device__ __forceinline__ float rcp_rn_normal(float s) {
    float r0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(s));
    const float e = fmaf(s, r0, -1.0f);
    return fmaf(r0, -e, r0);
}
// "regular" on fp16 bits: 0x0400 <= bits <= 0x7BFF (normal, finite, positive)
__device__ __forceinline__ bool scale_bits_regular(uint32_t sb) { return sb - 0x0400u < 0x7800u; }

// absmax of NW packed words as an fp16 bit pattern.  Integer max on |bits|: NaN patterns (> 0x7C00)
// are the largest, so a NaN anywhere propagates exactly like torch's abs().max().
template <int NW>
__device__ __forceinline__ uint32_t absmax_bits(const uint32_t (&p)[NW]) {
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < NW; ++i) m = __vmaxu2(m, p[i] & 0x7FFF7FFFu);
    return max(m & 0xffffu, m >> 16);
}

// Fake-quantize one group held as 2*NW halves per lane (NW packed words), LPG lanes per group.
// Returns true when the group was handled (regular scale).  Otherwise p is untouched and the caller
// runs literal_sym_h16 on the group's memory: keeping the rare literal path out of line (and out of
// the register tile) keeps the hot loop small.
template <int FMT, int LPG, int NW>
__device__ __forceinline__ bool sym_quant_tile_h16(uint32_t (&p)[NW], float& s, float delta) {
    using HG = typename SymFmt<FMT>::HG;
    uint32_t m = absmax_bits<NW>(p);
#pragma unroll
    for (int o = LPG / 2; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float a = h2f(uint16_t(m));
    const __half sh = scale_from_absmax_h16<HG>(a);
    s = __half2float(sh);
    if (!scale_bits_regular(__half_as_ushort(sh))) {
        s = rnd_in<__half>(__fdiv_rn(a, HG::VMAX));      // the literal path gets the literal scale (exact ties among subnormals)
        return false;
    }
    const float r = rcp_rn_normal(s);
    const uint64_t r2 = pk(r, r), s2 = pk(s, s);
#pragma unroll
    for (int i = 0; i < NW; ++i) p[i] = sym_pair_h16<HG>(p[i], r2, s2, delta);
    return true;
}

// Literal reference sequence for the elements of one lane, memory to memory (zero, subnormal, inf
// or NaN scale).  The lane's elements are `nv` vectors of `vec` halves at (j*lpg + lig)*vec.
static __device__ __noinline__ void literal_sym_h16(const __half* src, __half* dst, int lig, int lpg, int vec, int nv, float s, int gt) {
    const GridTable& g = c_grids[gt];
#pragma unroll 1
    for (int j = 0; j < nv; ++j) {
#pragma unroll 1
        for (int e = 0; e < vec; ++e) {
            const int idx = (j * lpg + lig) * vec + e;
            const float x = __half2float(src[idx]);
            dst[idx] = __float2half_rn(quant_elem_literal<__half, TIE_KERNEL>(x, s, g) * s);
        }
    }
}

static __device__ __noinline__ void literal_split_h16(const __half* src, __half* dst, int lig, int lpg, int vec, int nv, float sn, float sp,
                                                      int gtn, int gtp) {
    const GridTable& gn = c_grids[gtn];
    const GridTable& gp = c_grids[gtp];
#pragma unroll 1
    for (int j = 0; j < nv; ++j) {
#pragma unroll 1
        for (int e = 0; e < vec; ++e) {
            const int idx = (j * lpg + lig) * vec + e;
            const float x = __half2float(src[idx]);
            const float xn = (x <= 0.0f) ? x : 0.0f, xp = (x > 0.0f) ? x : 0.0f;          // quant_utils.py:428-429
            const float qn = scan_kernel_rule(rnd_in<__half>(__fdiv_rn(xn, sn)), gn.v, gn.k);
            const float qp = scan_kernel_rule(rnd_in<__half>(__fdiv_rn(xp, sp)), gp.v, gp.k);
            dst[idx] = __float2half_rn(__fadd_rn(__fmul_rn(qn, sn), __fmul_rn(qp, sp)));    // quant_utils.py:450
        }
    }
}

// A group that holds a NaN: where(x<=0, x, 0) / where(x>0, x, 0) turn the NaN into 0 on both sides
// (quant_utils.py:428-429), so the maxima ignore it.  Rare: every lane rescans the whole group.
static __device__ __noinline__ void literal_split_nan_group_h16(const __half* src, __half* dst, int lig, int lpg, int vec, int nv,
                                                                float nmax, float pmax, int gtn, int gtp) {
    float an = 0.0f, ap = 0.0f;
#pragma unroll 1
    for (int i = 0; i < 128; ++i) {
        const float f = __half2float(src[i]);
        an = fmaxf(an, (f <= 0.0f) ? -f : 0.0f);
        ap = fmaxf(ap, (f > 0.0f) ? f : 0.0f);
    }
    const float sn = rnd_in<__half>(__fdiv_rn(an, nmax));
    const float sp = rnd_in<__half>(__fdiv_rn(ap, pmax));
    literal_split_h16(src, dst, lig, lpg, vec, nv, sn, sp, gtn, gtp);
}

// ---- sign-split ----------------------------------------------------------------------------
// One pair: elements > 0 use (rp, sp, POS grid), elements < 0 use (rn, sn, NEG grid); +0 / -0 give +0
// on either side, so the sign bit alone picks the side.  No NaN can reach this function (groups that
// hold a NaN take the literal path).
// How many of the per-element side selects (r, s, magic scale[, sign mask]) run as integer
// multiply-adds on the FMA pipe instead of LOP3 on the ALU pipe.  Measured on B200 (tools/kbench.py,
// fc2 shape 25600 x 7680): 0 -> 5.40, 1 -> 5.70, 2 -> 5.92, 3 -> 6.15 TB/s.
#ifndef FPQ_SPLIT_IMAD
#define FPQ_SPLIT_IMAD 3
#endif
// bit-pattern select between a "positive side" and a "negative side" constant with m = -1 (negative
// element) or 0.  Two flavours so that the work can be spread over both math pipes of the SM
// sub-partition (ncu: the ALU pipe, where FSEL/LOP3/FMNMX/F2FP live, is the busy one): an integer
// multiply-add on the FMA pipe, or a LOP3 on the ALU pipe.
__device__ __forceinline__ float sel_imad(int m, float pos, float neg) {
    int d;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(m), "r"(__float_as_int(pos) - __float_as_int(neg)), "r"(__float_as_int(pos)));
    return __int_as_float(d);
}
__device__ __forceinline__ float sel_lop(int m, float pos, float neg) {
    return __int_as_float(__float_as_int(pos) ^ ((__float_as_int(pos) ^ __float_as_int(neg)) & m));
}

// Sides with a UNIFORM negative grid (e1m2: step 1/4 up to 1.75; int: step 1 up to 32) share the
// positive side's rounding constants: the negative side is rescaled by the power of two K that maps
// its step onto the positive format's subnormal step 2^(EMIN-M) (rn' = rn*K, sn' = sn/K, both exact),
// and p = 2^exponent(max(w, 2^EMIN)) is taken WITHOUT the absolute value, so every negative w gets the
// constant p = 2^EMIN -- exactly the uniform grid.  Only r and s are then selected per element.
template <class NEG, class POS> struct SplitScale {
    // NEG uniform <=> all of its values lie in its own subnormal region or first binade with the same step
    static constexpr bool UNIFORM_NEG = (NEG::M - NEG::EMIN == 2 && NEG::VNUM * 4 == 7 * NEG::VDEN) || (NEG::M == NEG::EMIN);
    // K = step(POS subnormal) / step(NEG)
    static constexpr float K = UNIFORM_NEG ? (Magic<POS>::EM / float(1u << POS::M)) / (Magic<NEG>::EM / float(1u << NEG::M)) : 1.0f;
};

template <class NEG, class POS>
__device__ __forceinline__ uint32_t split_pair_h16_w(float2 x, float rn, float sn, float rp, float sp, float delta);

template <class NEG, class POS>
__device__ __forceinline__ uint32_t split_pair_h16(uint32_t x2, float rn, float sn, float rp, float sp, float delta) {
    return split_pair_h16_w<NEG, POS>(__half22float2(*reinterpret_cast<const __half2*>(&x2)), rn, sn, rp, sp, delta);
}

template <class NEG, class POS>
__device__ __forceinline__ uint32_t split_pair_h16_w(float2 x, float rn, float sn, float rp, float sp, float delta) {
    // rn, sn arrive pre-scaled by the caller when SplitScale::UNIFORM_NEG (rn*K, sn/K)
#if FPQ_SPLIT_IMAD >= 4
    const int m0 = -int(__umulhi(__float_as_uint(x.x), 2u)), m1 = -int(__umulhi(__float_as_uint(x.y), 2u));
#else
    const int m0 = __float_as_int(x.x) >> 31, m1 = __float_as_int(x.y) >> 31;      // -1: negative side
#endif
#if FPQ_SPLIT_IMAD >= 1
    const float r0 = sel_imad(m0, rp, rn), r1 = sel_imad(m1, rp, rn);
#else
    const float r0 = sel_lop(m0, rp, rn), r1 = sel_lop(m1, rp, rn);
#endif
#if FPQ_SPLIT_IMAD >= 2
    const float s0 = sel_imad(m0, sp, sn), s1 = sel_imad(m1, sp, sn);
#else
    const float s0 = sel_lop(m0, sp, sn), s1 = sel_lop(m1, sp, sn);
#endif
    const uint32_t v2 = pack_h2_u64(fmul2(pk(x.x, x.y), pk(r0, r1)));
    uint64_t q;
    if constexpr (SplitScale<NEG, POS>::UNIFORM_NEG) {
        const float w0 = fhadd(uint16_t(v2 & 0xffffu), delta);
        const float w1 = fhadd(uint16_t(v2 >> 16), delta);
        const float p0 = __uint_as_float(__float_as_uint(fmaxf(w0, Magic<POS>::EM)) & 0x7F800000u);     // signed max
        const float p1 = __uint_as_float(__float_as_uint(fmaxf(w1, Magic<POS>::EM)) & 0x7F800000u);
        const uint64_t p = pk(p0, p1), w = pk(w0, w1);
        const uint64_t y = ffma2(p, pk(Magic<POS>::SC, Magic<POS>::SC), w);
        q = ffma2(p, pk(-Magic<POS>::SC, -Magic<POS>::SC), y);
    } else {
        static_assert(Magic<NEG>::EM == Magic<POS>::EM && Magic<NEG>::SC == Magic<POS>::SC, "non-uniform negative grids must share the positive format");
        q = round_pair_magic(v2, delta, Magic<POS>::EM, Magic<POS>::EM, Magic<POS>::SC, Magic<POS>::SC);
    }
    return pack_h2_u64(fmul2(q, pk(s0, s1)));
}

}  // namespace fpq
