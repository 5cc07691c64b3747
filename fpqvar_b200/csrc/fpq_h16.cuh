// Packed fast path for fp16 tensors under the kernel tie rule (the activation flow of the
// rotated models: fp16 in, fp16 out).
//
// The generic path (fpq_common.cuh) spends ~20-38 instructions per element and is issue-bound on
// B200 (ncu: 85 % issue-slot utilisation at 36 % DRAM, profiles/r1a_generic_kernels_ncu.txt).  At 4 bytes per
// element the HBM roofline leaves ~20 issue slots per element at full clocks and ~16 under the 1000 W
// cap, so this path is written for instruction count (~5 per element for the hardware formats):
//   * two elements per instruction wherever the ISA allows it: FMUL2 / FFMA2 (packed fp32,
//     sm_100), F2FP.PACK_AB, HMNMX2 on |x| for the absmax, HFMA2 for the rescale;
//   * grid rounding by the FP4 / FP6 CONVERSION HARDWARE of sm_100a for the formats it knows
//     (e2m1, e2m3, e3m2 = F2FP.SATFINITE.E2M1 / .E2M3 / .E3M2 and the UNPACK_B back to fp16):
//       v = half(x * RN(1/s)) == half(x/s)   (proof: fpq_common.cuh)
//       w = v + 2^-17          widened by the mixed-precision add FHADD (add.f32.f16), which adds the
//                              tie shift for free: +2^-17 moves exact midpoints to the side the kernel
//                              rule sends them to (+inf) and is smaller than the distance from any
//                              other fp16 value to a midpoint, so nothing else changes side; v has 11
//                              significant bits, so v + 2^-17 is exact in fp32 and never a tie.
//       q = fp16( cvt.rn.satfinite.<fmt>x2( w ) )          two elements per instruction, both ways
//       out = fma.rn.f16x2(q, s, +0)   q has <= 4 significant bits and s 11, so the fp16 FMA rounds the
//                              EXACT product once, like half(float(q) * float(s)); the +0 addend turns
//                              the -0 the sign-magnitude formats return for tiny negative w into the
//                              reference's +0 (its grid holds +0.0 only);
//   * for the other formats (e1m2, e3m0) the same w goes through a magic-number FFMA:
//       p = 2^max(exponent(w), EMIN);  y = RN(p * 1.5*2^(23-M) + w);  q = y - p * 1.5*2^(23-M)
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
__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c) {     // fp16x2 fused multiply-add, one rounding
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t dup_h(__half h) { return uint32_t(__half_as_ushort(h)) * 0x00010001u; }

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

// ---- grid rounding of a packed pair ------------------------------------------------------------
// Formats the sm_100a conversion hardware knows: fp32 pair -> two 4/6-bit codes -> fp16 pair, two
// instructions (F2FP.SATFINITE.<fmt>.F32.PACK_AB_MERGE_C + F2FP.F16.<fmt>.UNPACK_B).  The hardware
// rounds to nearest EVEN; the caller's tie shift has removed every tie, so "nearest" is all that is
// used.  satfinite clamps to +-VMAX, which is also what the reference's scan answers above the grid.
// PRE: e1m2 (uniform, step 1/4 up to 1.75) is not a hardware format, but HALF of it is the uniform low end of e2m3
// (step 1/8 up to 0.875): the group's reciprocal scale is halved and its scale doubled, both exact.  The doubled
// scale must stay a finite fp16 number: S_MAX_BITS bounds the "regular" scales of the format.
template <class HG> struct HwCvt { static constexpr bool AVAILABLE = false; static constexpr float PRE = 1.0f; static constexpr uint32_t S_MAX_BITS = 0x7BFFu; };
template <> struct HwCvt<HG_E2M1> {
    static constexpr bool AVAILABLE = true;
    static constexpr float PRE = 1.0f;
    static constexpr uint32_t S_MAX_BITS = 0x7BFFu;
    static __device__ __forceinline__ uint32_t round_trip(float w0, float w1) {
        uint32_t q2;
        asm("{ .reg .b8 t; cvt.rn.satfinite.e2m1x2.f32 t, %1, %2; cvt.rn.f16x2.e2m1x2 %0, t; }" : "=r"(q2) : "f"(w1), "f"(w0));
        return q2;
    }
};
template <> struct HwCvt<HG_E2M3> {
    static constexpr bool AVAILABLE = true;
    static constexpr float PRE = 1.0f;
    static constexpr uint32_t S_MAX_BITS = 0x7BFFu;
    static __device__ __forceinline__ uint32_t round_trip(float w0, float w1) {
        uint32_t q2;
        asm("{ .reg .b16 t; cvt.rn.satfinite.e2m3x2.f32 t, %1, %2; cvt.rn.f16x2.e2m3x2 %0, t; }" : "=r"(q2) : "f"(w1), "f"(w0));
        return q2;
    }
};
template <> struct HwCvt<HG_E3M2> {
    static constexpr bool AVAILABLE = true;
    static constexpr float PRE = 1.0f;
    static constexpr uint32_t S_MAX_BITS = 0x7BFFu;
    static __device__ __forceinline__ uint32_t round_trip(float w0, float w1) {
        uint32_t q2;
        asm("{ .reg .b16 t; cvt.rn.satfinite.e3m2x2.f32 t, %1, %2; cvt.rn.f16x2.e3m2x2 %0, t; }" : "=r"(q2) : "f"(w1), "f"(w0));
        return q2;
    }
};
template <> struct HwCvt<HG_E1M2> {
    static constexpr bool AVAILABLE = true;
    static constexpr float PRE = 0.5f;
    static constexpr uint32_t S_MAX_BITS = 0x77FFu;          // 2 * s stays finite (callers that use the conversion hardware check it)
    static __device__ __forceinline__ uint32_t round_trip(float w0, float w1) { return HwCvt<HG_E2M3>::round_trip(w0, w1); }
};

// Any other half grid: q = R_K(w) by a magic-number FFMA, as packed fp32.
__device__ __forceinline__ uint64_t round_pair_magic(float w0, float w1, float em, float sc) {
    const float p0 = __uint_as_float(__float_as_uint(fmaxf(fabsf(w0), em)) & 0x7F800000u);
    const float p1 = __uint_as_float(__float_as_uint(fmaxf(fabsf(w1), em)) & 0x7F800000u);
    const uint64_t p = pk(p0, p1), w = pk(w0, w1);
    const uint64_t y = ffma2(p, pk(sc, sc), w);
    return ffma2(p, pk(-sc, -sc), y);
}

// Per-group constants of the symmetric flow of the quantizer kernels: r = RN(1/s) and s as packed fp32 pairs.
struct SymK {
    uint64_t r2, s2;
};
template <class HG>
__device__ __forceinline__ SymK make_symk(float s, float r) {
    SymK k;
    k.r2 = pk(r, r);
    k.s2 = pk(s, s);
    return k;
}

// The symmetric element function on the conversion hardware (formats with HwCvt<HG>::AVAILABLE): the format scorer's
// (fpq_score.cu); exhaustively checked against the literal sequence by fpq_selftest_f16_flow(32 + format).
// r2 = RN(1/s) * PRE, sh2 = fp16(s / PRE) twice.  The quantizer kernels do NOT use it: measured at equal conditions
// (profiles/r2_quantizer_rounding_ab.txt) it is no faster there -- the pack runs at a quarter of the issue rate.
template <class HG>
__device__ __forceinline__ uint32_t sym_pair_h16_hw(uint64_t xf2, uint64_t r2, uint32_t sh2, float delta) {
    const uint32_t v2 = pack_h2_u64(fmul2(xf2, r2));                              // half(x/s) (times PRE, exact)
    const uint32_t q2 = HwCvt<HG>::round_trip(fhadd(uint16_t(v2 & 0xffffu), delta), fhadd(uint16_t(v2 >> 16), delta));
    return hfma2(q2, sh2, 0u);                                                    // half(q*s), -0 -> +0
}
template <class HG> __device__ __forceinline__ bool scale_bits_regular_hw(uint32_t sb) { return sb - 0x0400u <= HwCvt<HG>::S_MAX_BITS - 0x0400u; }

// One packed pair of the symmetric flow: two fp16 inputs (already widened to packed fp32) -> two fp16
// outputs q*s.
template <class HG>
__device__ __forceinline__ uint32_t sym_pair_h16_w(uint64_t xf2, const SymK& k, float delta) {
    const uint32_t v2 = pack_h2_u64(fmul2(xf2, k.r2));                            // half(x/s)
    const uint64_t q = round_pair_magic(fhadd(uint16_t(v2 & 0xffffu), delta), fhadd(uint16_t(v2 >> 16), delta), Magic<HG>::EM, Magic<HG>::SC);
    return pack_h2_u64(fmul2(q, k.s2));                                           // half(q*s)
}
template <class HG>
__device__ __forceinline__ uint32_t sym_pair_h16(uint32_t x2, const SymK& k, float delta) {
    return sym_pair_h16_w<HG>(widen_h2(x2), k, delta);
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
template <class HG> __device__ __forceinline__ bool scale_bits_regular_for(uint32_t sb) { return scale_bits_regular(sb); }

// absmax of NW packed words as an fp16 bit pattern, one HMNMX2.NAN on |a|, |b| per word.  A NaN anywhere
// comes out as a NaN pattern (> 0x7C00), exactly like torch's abs().max().
__device__ __forceinline__ uint32_t hmax2_abs_nan(uint32_t a, uint32_t b) {
    const __half2 r = __hmax2_nan(__habs2(*reinterpret_cast<const __half2*>(&a)), __habs2(*reinterpret_cast<const __half2*>(&b)));
    return *reinterpret_cast<const uint32_t*>(&r);
}
template <int NW>
__device__ __forceinline__ uint32_t absmax_bits(const uint32_t (&p)[NW]) {
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < NW; ++i) m = hmax2_abs_nan(m, p[i]);
    return max(m & 0xffffu, m >> 16);
}

// Fake-quantize one group held as 2*NW halves per lane (NW packed words), LPG lanes per group.
// Returns true when the group was handled (regular scale).  Otherwise p is untouched and the caller
// runs literal_sym_h16 on the group's memory: keeping the rare literal path out of line (and out of
// the register tile) keeps the hot loop small.
// HW = true: the element function on the FP4 / FP6 conversion hardware (formats that have it): 5 instructions per pair
// instead of 10 after the division.  Chosen per kernel from measurements: the register-tile kernels are no faster with it
// (profiles/r2_quantizer_rounding_ab.txt); the streaming rotate kernel, which is issue-bound, is.
template <int FMT, int LPG, int NW, bool HW = true>
__device__ __forceinline__ bool sym_quant_tile_h16(uint32_t (&p)[NW], float& s, float delta) {
    using HG = typename SymFmt<FMT>::HG;
    if constexpr (HW && HwCvt<HG>::AVAILABLE) {
        uint32_t m = absmax_bits<NW>(p);
#pragma unroll
        for (int o = LPG / 2; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        const float a = h2f(uint16_t(m));
        const __half sh = scale_from_absmax_h16<HG>(a);
        s = __half2float(sh);
        if (!scale_bits_regular_hw<HG>(__half_as_ushort(sh))) {
            s = rnd_in<__half>(__fdiv_rn(a, HG::VMAX));
            return false;
        }
        const float rr = rcp_rn_normal(s) * HwCvt<HG>::PRE;
        const uint64_t r2 = pk(rr, rr);
        const uint32_t sh2 = dup_h(__float2half_rn(s * (1.0f / HwCvt<HG>::PRE)));
#pragma unroll
        for (int i = 0; i < NW; ++i) p[i] = sym_pair_h16_hw<HG>(widen_h2(p[i]), r2, sh2, delta);
        return true;
    }
    uint32_t m = absmax_bits<NW>(p);
#pragma unroll
    for (int o = LPG / 2; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float a = h2f(uint16_t(m));
    const __half sh = scale_from_absmax_h16<HG>(a);
    s = __half2float(sh);
    if (!scale_bits_regular_for<HG>(__half_as_ushort(sh))) {
        s = rnd_in<__half>(__fdiv_rn(a, HG::VMAX));      // the literal path gets the literal scale (exact ties among subnormals)
        return false;
    }
    const SymK k = make_symk<HG>(s, rcp_rn_normal(s));
#pragma unroll
    for (int i = 0; i < NW; ++i) p[i] = sym_pair_h16<HG>(p[i], k, delta);
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
// Elements > 0 use (rp, sp, POS grid), elements < 0 use (rn, sn, NEG grid); +0 / -0 give +0 on either side, so the sign bit
// alone picks the side.  No NaN can reach these functions (groups that hold a NaN take the literal path).
// BOTH sides are evaluated for the whole pair with packed instructions and uniform constants, and the sign bits pick the
// result per half at the end:
//   positive side   w = half(x * rp) + 2^-17 -> conversion hardware (e2m1 / e2m3), as in sym_pair_h16_hw
//   negative side   a UNIFORM grid (e1m2: step 1/4 up to 1.75; int: step 1 up to 32) is rounded entirely in packed fp16 with
//                   a magic constant C = 1.5 * 2^10 * step, whose ulp is the step:  y = fma.rn.f16x2(v, 1 - 2^-11, C);
//                   q = y - C.  The factor 1 - 2^-11 moves v towards zero by less than one fp16 ulp: exact midpoints fall
//                   to the side the kernel rule sends them to (the larger value, i.e. towards zero for v <= 0) and no other
//                   fp16 value reaches or crosses a midpoint.  A non-uniform negative grid (afpq: e2m1 on both sides) goes
//                   through the conversion hardware too.
// A lane of the wrong side computes garbage (possibly inf / NaN) that the select discards.
// Measured against the round-1 element function (per-element constants selected with integer multiply-adds, magic-number
// rounding on both sides; profiles/r2_signsplit_ab.txt): 10 % slower on one isolated launch at full clocks (5.73 vs 6.36 TB/s),
// 7 % faster sustained under the 1000 W cap (6.07 vs 5.68 TB/s at 487 vs 592 W), 8 % faster inside the step (5.59 vs 5.18).
template <class NEG> struct UniformNeg {
    static constexpr bool OK = (NEG::M - NEG::EMIN == 2 && NEG::VNUM * 4 == 7 * NEG::VDEN) || (NEG::M == NEG::EMIN);
    static constexpr float STEP = Magic<NEG>::EM / float(1u << NEG::M);
    static constexpr float C = 1536.0f * STEP;                 // 1.5 * 2^10 * step
};
struct SplitK {
    uint64_t rn2, rp2;          // RN(1/s) of each side as packed fp32 pairs (0 for a side without elements)
    uint32_t snh2, sph2;        // the two scales as fp16 pairs
};
template <class NEG, class POS>
__device__ __forceinline__ SplitK make_splitk(float sn, float rn, float sp, float rp) {
    SplitK k;
    k.rn2 = pk(rn, rn);
    k.rp2 = pk(rp, rp);
    k.snh2 = dup_h(__float2half_rn(sn));
    k.sph2 = dup_h(__float2half_rn(sp));
    return k;
}
template <class NEG, class POS>
__device__ __forceinline__ uint32_t split_pair_h16_w(uint64_t xf2, uint32_t x2, const SplitK& k, float delta) {
    static_assert(HwCvt<POS>::AVAILABLE && HwCvt<POS>::PRE == 1.0f, "the positive side of every sign-split format is e2m1 or e2m3");
    const uint32_t vp2 = pack_h2_u64(fmul2(xf2, k.rp2));                           // half(x/sp)
    const uint32_t vn2 = pack_h2_u64(fmul2(xf2, k.rn2));                           // half(x/sn)
    const uint32_t qp2 = HwCvt<POS>::round_trip(fhadd(uint16_t(vp2 & 0xffffu), delta), fhadd(uint16_t(vp2 >> 16), delta));
    uint32_t qn2;
    if constexpr (UniformNeg<NEG>::OK) {
        const uint32_t c2 = dup_h(__float2half_rn(UniformNeg<NEG>::C));
        qn2 = hadd2(hfma2(vn2, 0x3BFF3BFFu, c2), c2 ^ 0x80008000u);                // (v * (1 - 2^-11) + C) - C
    } else {
        static_assert(HwCvt<NEG>::AVAILABLE && HwCvt<NEG>::PRE == 1.0f, "non-uniform negative grids must be hardware formats");
        qn2 = HwCvt<NEG>::round_trip(fhadd(uint16_t(vn2 & 0xffffu), delta), fhadd(uint16_t(vn2 >> 16), delta));
    }
    uint32_t neg;                                                                  // 0xFFFF in every half whose sign bit is set
    asm("prmt.b32 %0, %1, 0, 0xBB99;" : "=r"(neg) : "r"(x2));
    const uint32_t q2 = (qp2 & ~neg) | (qn2 & neg);
    const uint32_t s2 = (k.sph2 & ~neg) | (k.snh2 & neg);
    return hfma2(q2, s2, 0u);                                                      // half(q*s), -0 -> +0
}

template <class NEG, class POS>
__device__ __forceinline__ uint32_t split_pair_h16(uint32_t x2, const SplitK& k, float delta) {
    return split_pair_h16_w<NEG, POS>(widen_h2(x2), x2, k, delta);
}

}  // namespace fpq
