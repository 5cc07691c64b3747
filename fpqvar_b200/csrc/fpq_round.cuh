// Grid formats and the element-wise rounding rules, device side.
//
// Two implementations of each rule live here on purpose:
//   * scan_kernel_rule / scan_argmin_rule: the literal reference loops
//     (quant/quant_kernel.cu:25-37 and quantize_to_nearest_grid, quant_utils.py:224-230).
//     Used for unknown grids, for the rare "irregular" groups, and as the comparator of the
//     exhaustive self-test.
//   * round_closed<HG, TIE>: branch-free closed form by integer manipulation of the fp32 bit
//     pattern, used on the hot path.  fpq_selftest_rounding proves it equal to the scan for
//     all 2^32 inputs of every format / tie rule.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace fpq {

enum : int { TIE_KERNEL = 0, TIE_ARGMIN = 1 };

// ---------------------------------------------------------------------------------------
// Half grids.  A half grid is the non-negative side of a low-bit float format:
//   values  k * 2^(EMIN-M)              for k = 0 .. 2^M        ("subnormal" region, < 2^EMIN)
//           (2^M + j) * 2^(e-M)         for e >= EMIN, j < 2^M  (normal binades)
//   capped at VMAX = VNUM / VDEN.
// A uniform integer grid {0..32} is the degenerate case EMIN = M = 5.
// ---------------------------------------------------------------------------------------
template <int EMIN_, int M_, int VNUM_, int VDEN_>
struct HalfGrid {
    static constexpr int EMIN = EMIN_;
    static constexpr int M = M_;
    static constexpr int VNUM = VNUM_, VDEN = VDEN_;
    static constexpr float VMAX = float(VNUM_) / float(VDEN_);
    static constexpr uint32_t LOW = (1u << (23 - M_)) - 1u;    // discarded mantissa bits
    static constexpr uint32_t HALF = 1u << (22 - M_);          // half of the kept LSB
    static constexpr uint32_t EM_BITS = uint32_t(127 + EMIN_) << 23;   // bits of 2^EMIN
    __host__ __device__ static constexpr float em() { return EMIN_ >= 0 ? float(1u << (EMIN_ >= 0 ? EMIN_ : 0)) : 1.0f / float(1u << (EMIN_ < 0 ? -EMIN_ : 0)); }
};

using HG_E2M1 = HalfGrid<0, 1, 6, 1>;
using HG_E1M2 = HalfGrid<0, 2, 7, 4>;
using HG_E3M0 = HalfGrid<-2, 0, 16, 1>;
using HG_E2M3 = HalfGrid<0, 3, 15, 2>;
using HG_E3M2 = HalfGrid<-2, 2, 28, 1>;
using HG_INT32 = HalfGrid<5, 5, 32, 1>;

enum : int { HGID_E2M1 = 0, HGID_E1M2 = 1, HGID_E3M0 = 2, HGID_E2M3 = 3, HGID_E3M2 = 4, HGID_INT32 = 5, HGID_COUNT = 6 };

// ---------------------------------------------------------------------------------------
// The reference's grids, spelled out (constant memory) for the literal scans.
// ---------------------------------------------------------------------------------------
struct GridTable {
    int k;
    float v[64];
};
// index: 0..4 symmetric FPQ_FMT_*, 5 int_neg, 6 e2m3_pos, 7 e1m2_neg, 8 e2m1_pos, 9 e2m1_neg
enum : int { GT_E2M1 = 0, GT_E1M2, GT_E3M0, GT_E2M3, GT_E3M2, GT_INT_NEG, GT_E2M3_POS, GT_E1M2_NEG, GT_E2M1_POS, GT_E2M1_NEG, GT_COUNT };

__device__ __forceinline__ float scan_kernel_rule(float x, const float* __restrict__ g, int k) {
    float best = 102400.0f, z = 0.0f;
    for (int i = 0; i < k; ++i) {
        float d = fabsf(x - g[i]);
        if (d <= best) { best = d; z = g[i]; }
    }
    return z;
}

__device__ __forceinline__ float scan_argmin_rule(float x, const float* __restrict__ g, int k) {
    float best = fabsf(x - g[0]);
    float z = g[0];
    for (int i = 1; i < k; ++i) {
        float d = fabsf(x - g[i]);
        if (!(best != best) && ((d != d) || d < best)) { best = d; z = g[i]; }
    }
    return z;
}

template <int TIE>
__device__ __forceinline__ float scan_rule(float x, const float* __restrict__ g, int k) {
    return TIE == TIE_KERNEL ? scan_kernel_rule(x, g, k) : scan_argmin_rule(x, g, k);
}

// ---------------------------------------------------------------------------------------
// Closed form, symmetric grid, FINITE input.
//   TIE_KERNEL: exact ties go to the larger value  (toward +inf)
//   TIE_ARGMIN: exact ties go to the smaller value (toward -inf)
// Values below 2^EMIN are moved into the first normal binade by adding +-2^EMIN with a
// DIRECTED rounding (toward the side ties go to), which keeps "below / on / above a
// midpoint" intact even when the addition has to drop low bits; the mantissa is then rounded
// by an integer add-and-mask, whose carry walks into the exponent for free.
// The result is +0 (never -0) when it rounds to zero, as in the reference (the grid's zero
// is +0.0).  CLAMP=false may be used when |v| is known to stay below the first midpoint
// above VMAX (true for in-group values: |x| <= absmax).
// ---------------------------------------------------------------------------------------
template <class HG, int TIE, bool CLAMP>
__device__ __forceinline__ float round_closed(float v) {
    const uint32_t vb = __float_as_uint(v);
    const bool sub = fabsf(v) < HG::em();
    const float off = sub ? __uint_as_float((vb & 0x80000000u) | HG::EM_BITS) : 0.0f;
    const float y = (TIE == TIE_KERNEL) ? __fadd_rd(v, off) : __fadd_ru(v, off);
    const uint32_t yb = __float_as_uint(y);
    const uint32_t neg = yb >> 31;
    uint32_t t = yb + HG::HALF - (TIE == TIE_KERNEL ? neg : 1u - neg);
    t &= ~HG::LOW;
    float q = __uint_as_float(t) - off;
    if (CLAMP) q = fminf(fmaxf(q, -HG::VMAX), HG::VMAX);
    return q;
}

// Same, but also returns whether v sits within `DELTA` fp32 ulps of a rounding boundary
// (used when v is only an approximation of the reference's value).
template <class HG, int TIE>
__device__ __forceinline__ float round_closed_near(float v, bool& near) {
    constexpr uint32_t DELTA = 32u;
    const uint32_t vb = __float_as_uint(v);
    const bool sub = fabsf(v) < HG::em();
    const float off = sub ? __uint_as_float((vb & 0x80000000u) | HG::EM_BITS) : 0.0f;
    const float y = (TIE == TIE_KERNEL) ? __fadd_rd(v, off) : __fadd_ru(v, off);
    const uint32_t yb = __float_as_uint(y);
    const uint32_t neg = yb >> 31;
    uint32_t t = yb + HG::HALF - (TIE == TIE_KERNEL ? neg : 1u - neg);
    near = ((t + DELTA) & HG::LOW) < 2u * DELTA;
    t &= ~HG::LOW;
    return __uint_as_float(t) - off;
}

// Full element rule for a symmetric grid on ANY fp32 input (NaN, inf, huge): what
// quant_cuda.quant / quantize_to_nearest_grid return for the reference's symmetric tables.
template <class HG, int TIE>
__device__ __forceinline__ float round_any_sym(float v) {
    if (TIE == TIE_KERNEL) {
        // the scan starts from best = 102400 and z = 0: nothing within 102400 -> +0
        // (NaN and inf compare false).  The nearest entry to a huge v is +-VMAX.
        const float d = fabsf(v - copysignf(HG::VMAX, v));
        if (!(d <= 102400.0f)) return 0.0f;
    } else {
        // argmin: a NaN distance wins at index 0; all-inf distances also give index 0
        if (!(fabsf(v) <= 3.4028234663852886e38f)) return -HG::VMAX;
    }
    return round_closed<HG, TIE, true>(v);
}

// One-sided grids of the sign-split formats: the negative grid holds {-VMAX..0}, the
// positive grid {0..VMAX}; an input on the wrong side is nearest to 0.
template <class HG, int TIE, bool NEGATIVE_SIDE>
__device__ __forceinline__ float round_any_onesided(float v) {
    if (TIE == TIE_KERNEL) {
        const float edge = NEGATIVE_SIDE ? (v < 0.f ? -HG::VMAX : 0.f) : (v > 0.f ? HG::VMAX : 0.f);
        const float d = fabsf(v - edge);
        if (!(d <= 102400.0f)) return 0.0f;
    } else {
        if (!(fabsf(v) <= 3.4028234663852886e38f)) return NEGATIVE_SIDE ? -HG::VMAX : 0.0f;
    }
    float q = round_closed<HG, TIE, true>(v);
    q = NEGATIVE_SIDE ? fminf(q, 0.0f) : fmaxf(q, 0.0f);
    return q + 0.0f;    // -0 -> +0 cannot occur (closed form yields +0), kept as documentation
}

}  // namespace fpq
