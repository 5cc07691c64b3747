// libfpq_b200 -- fused GALT transform + 128-block random-Hadamard rotation (+ fake quant).
// Replaces (reference rows a8, a9 of SURVEY.md section 8):
//   activations  basic_var.py:263,266   matmul(x.mul(s), Q)  -> QuantizedLinear.forward act_quant
//   weights      transform_model_utils.py:8-28 (W / s) then rotation_utils.py:129-154 (W.double() @ Q)
// Q = I (x) diag(sigma) H_128 / fl32(sqrt(128)) (rotation_utils.py:69-104, hadamard_utils.py:63-99),
// so (x @ Q) on an aligned 128-chunk is FWHT_128(x * sigma) / c: 7 butterfly stages instead of a
// dense C x C GEMM.
//
// Register tile: the same 8-lanes-per-128-group / 16-values-per-lane layout as the quant kernels
// (fpq_common.cuh).  With fp32 input, value v[4j+k] of lane l is element e = 32j + 4l + k of the
// chunk, so the seven index bits split as k (2 bits, in registers), l (3 bits, across lanes:
// shfl.xor 1,2,4) and j (2 bits, in registers).  A Sylvester Hadamard transform is the tensor
// product of a 2-point butterfly over every index bit, in any order (here: 0, 1, 4, 5, 6, 2, 3, in both layouts).
#include <cstdlib>
#include "fpq_h16.cuh"

#ifndef FPQ_ROT_V2
#define FPQ_ROT_V2 1
#endif

namespace fpq {

struct SignMask { uint32_t w[4]; };    // bit e of the 128-bit mask set  <=>  sigma[e] = +1

// adaLN modulate operands (fpq_modulate_transform_rotate_quant): t = (x * (scale[b, c] + 1) + shift[b, c]) * smooth[c]
struct Modulate { const float* scale; const float* shift; size_t rows_per_batch; float one; };   // one: 1.0f, or -0.0f (the exact additive identity) under FPQ_MOD_GAIN

// 1 / fl32(sqrt(128)), rounded to fp32 (0x3DB504F3)
__device__ __forceinline__ float inv_sqrt128() { return __uint_as_float(0x3DB504F3u); }

// Activation kernel.  Work mapping: every 8-lane set owns ONE column chunk cc (so its 16
// smooth*sign multipliers are loaded once and live in registers) and walks down the rows
// k, k+K, k+2K, ... of that chunk column.  Arithmetic is two elements per instruction where the
// ISA allows (FMUL2/FFMA2/FADD2 packed fp32, sm_100): P[2j] = (v[4j], v[4j+1]), P[2j+1] = (v[4j+2], v[4j+3]).
// MOD (adaLN modulate fused in): a lane set walks CONTIGUOUS rows (k0*R .. k0*R + R - 1) instead of strided ones, so the
// (scale+1) and shift values of its column change only when the batch index does and live in registers in between --
// read from L2 once per (batch, column, row range) instead of once per chunk.
// Occupancy hint of the adaLN-fused variant (MOD).  Compiled freely it takes 100 registers = 2 CTAs per SM and is latency-bound
// (ncu r1e: 24 % warps active, long_scoreboard 2.5 per issue).  Measured (tools/kbench.py, 102400 x 1920, same run A/B,
// profiles/r1_kbench.txt): plain bounds 4314 / 4317 GB/s; declared (256, 1) -- still 2 CTAs, but the compiler spends 114
// registers and spills nothing -- 4526 / 4530; forced to 3 CTAs (80 registers, 48 B of spills) 3917; 4 CTAs (64 registers,
// 120 B) 4338.  So MOD is declared (256, 1).  The plain variants keep the plain bounds: the (256, 1) form costs them
// registers (67 -> 76) for nothing measured.
#ifndef FPQ_ROT_MOD_CTAS
#define FPQ_ROT_MOD_CTAS 1
#endif
#if FPQ_ROT_MOD_CTAS >= 1
#define FPQ_ROT_BOUNDS __launch_bounds__(256, MOD ? FPQ_ROT_MOD_CTAS : 0)      // 0 = no occupancy hint
#else
#define FPQ_ROT_BOUNDS __launch_bounds__(256)
#endif
template <int FMT, bool QUANT, bool MOD>
__global__ void FPQ_ROT_BOUNDS transform_rotate_quant_kernel(const float* __restrict__ x, const float* __restrict__ smooth,
                                                                     SignMask sm, __half* __restrict__ out, __half* __restrict__ rotated,
                                                                     size_t n_rows, int cpr, size_t sets_per_col, Modulate mod) {
    constexpr int LPG = 8;
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int lig = lane % LPG;
    const size_t ls = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) / LPG;       // lane-set id
    const int cc = int(ls % size_t(cpr));
    const size_t k0 = ls / size_t(cpr);
    const bool active = k0 < sets_per_col;
    const size_t row_stride = size_t(cpr) * 128;

    // multipliers m = smooth * sign for the 16 chunk positions of this lane (sign flips are exact,
    // so (x*s)*sigma == x*(s*sigma) bit for bit)
    uint64_t ms[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 s4 = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
        if (smooth != nullptr) s4 = __ldg(reinterpret_cast<const float4*>(smooth + size_t(cc) * 128 + (j * LPG + lig) * 4));
        float f[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (!((sm.w[j] >> (4 * lig + k)) & 1u)) f[k] = -f[k];
        ms[2 * j] = pk(f[0], f[1]);
        ms[2 * j + 1] = pk(f[2], f[3]);
    }
    const uint64_t neg1 = pk(-1.0f, -1.0f);
    const float delta = tie_delta_kernel(uint32_t(ls >> 36));
    const float cinv = inv_sqrt128();
    const uint64_t cinv2 = pk(cinv, cinv);

    // uniform trip count across the warp (the shuffles need every lane)
    const size_t trips = (n_rows + sets_per_col - 1) / sets_per_col;
    uint64_t A[MOD ? 8 : 1], SH[MOD ? 8 : 1];         // (scale + 1) and shift of this lane's 16 columns for batch cur_b
    size_t cur_b = ~size_t(0);
    pdl_wait();                                       // smooth / sign mask are parameters; x may come from the previous kernel
    for (size_t t = 0; t < trips; ++t) {
        const size_t row = MOD ? k0 * trips + t : k0 + t * sets_per_col;
        const bool valid = active && row < n_rows;
        const size_t off = row * row_stride + size_t(cc) * 128;
        uint64_t P[8];
        if (valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint4 u = ldg_stream(x + off + (j * LPG + lig) * 4);
                P[2 * j] = (uint64_t(u.y) << 32) | u.x;
                P[2 * j + 1] = (uint64_t(u.w) << 32) | u.z;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) P[i] = 0ull;
        }
        if constexpr (MOD) {
            const size_t b = valid ? row / mod.rows_per_batch : cur_b;
            if (b != cur_b) {
                cur_b = b;
                const size_t mo = b * row_stride + size_t(cc) * 128;
                const uint64_t one2 = pk(mod.one, mod.one);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 sc = __ldg(reinterpret_cast<const float4*>(mod.scale + mo + (j * LPG + lig) * 4));
                    const float4 sh = __ldg(reinterpret_cast<const float4*>(mod.shift + mo + (j * LPG + lig) * 4));
                    A[2 * j] = fadd2(pk(sc.x, sc.y), one2);               // scale.add(1)
                    A[2 * j + 1] = fadd2(pk(sc.z, sc.w), one2);
                    SH[2 * j] = pk(sh.x, sh.y);
                    SH[2 * j + 1] = pk(sh.z, sh.w);
                }
            }
            // .mul(scale.add(1)).add_(shift): the product with the SCALAR mul.rn.f32 (never contracted; ptxas fuses
            // mul.rn.f32x2 + add.rn.f32x2 into one FFMA2, which would skip the rounding of the product)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const F2 xv = unpk(P[i]), av = unpk(A[i]);
                P[i] = fadd2(pk(__fmul_rn(xv.lo, av.lo), __fmul_rn(xv.hi, av.hi)), SH[i]);
            }
        }
        // x * (s * sigma): basic_var.py:263 `.mul(s)` in fp32, then the sign row of Q
#pragma unroll
        for (int i = 0; i < 8; ++i) P[i] = fmul2(P[i], ms[i]);
        // index bit 0: inside a packed pair
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const F2 f = unpk(P[i]);
            P[i] = pk(f.lo + f.hi, f.lo - f.hi);
        }
        // The butterfly ORDER is part of the numerics (it fixes the fp32 summation tree).  Both layouts use the
        // same one -- index bits 0, 1, 4, 5, 6, 2, 3 -- so a row rotates to the same bits whichever kernel the
        // launcher picks for the tensor's size.
        auto reg_stage = [&](int h) {                 // between packed registers at distance h in P[]
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if ((i & h) == 0) {
                    const uint64_t a = P[i], b = P[i + h];
                    P[i] = fadd2(a, b);
                    P[i + h] = ffma2(b, neg1, a);
                }
            }
        };
        auto lane_stage = [&](int o) {                // across the lanes of the set
            const float sg = (lig & o) ? -1.0f : 1.0f;
            const uint64_t sg2 = pk(sg, sg);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const F2 f = unpk(P[i]);
                const uint64_t q = pk(__shfl_xor_sync(0xffffffffu, f.lo, o), __shfl_xor_sync(0xffffffffu, f.hi, o));
                P[i] = ffma2(P[i], sg2, q);            // upper lane: partner - mine ; lower lane: mine + partner (exact: * +-1)
            }
        };
        reg_stage(1);      // index bit 1
        lane_stage(4);     // index bit 4
        reg_stage(2);      // index bit 5
        reg_stage(4);      // index bit 6
        lane_stage(1);     // index bit 2
        lane_stage(2);     // index bit 3
        // / fl32(sqrt(128)), rounded to fp16: the fp16 GEMM output of the reference
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = pack_h2_u64(fmul2(P[i], cinv2));
        if (valid && rotated != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) stg_stream(rotated + off + (j * LPG + lig) * 4, make_uint2(w[2 * j], w[2 * j + 1]));
        }
        bool ok = true;
        float s = 0.0f;
        if constexpr (QUANT) ok = sym_quant_tile_h16<FMT, LPG, 8>(w, s, delta);
        if (valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) stg_stream(out + off + (j * LPG + lig) * 4, make_uint2(w[2 * j], w[2 * j + 1]));
            // irregular scale (zero / subnormal / inf / NaN): `out` now holds this lane's rotated values;
            // quantize them in place with the literal reference sequence
            if (!ok) literal_sym_h16(out + off, out + off, lig, LPG, 4, 4, s, SymFmt<FMT>::GT);
        }
    }
}

// Activation kernel, second layout (default): 4 lanes per chunk, 32 values (16 packed registers) per
// lane: vector j (0..7) of lane l holds elements 16j + 4l + k, so only index bits 2 and 3 cross lanes
// (2 SHFL per element instead of 3, and half the per-group scalar work per element).  The
// smooth*sign multipliers of the whole row live in shared memory (n_cols floats), which frees the
// registers the first layout spent on them and lets the chunks be walked with a plain grid stride.
// MOD: the adaLN modulate in front of it is fused as well (basic_var.py:263,266):
//     t = (x * (scale[b, c] + 1) + shift[b, c]) * smooth[c]        b = row / rows_per_batch
// as four separately rounded fp32 operations, exactly the reference's `.mul(scale.add(1)).add_(shift).mul(s)`.

template <int FMT, bool QUANT, bool MOD>
__global__ void __launch_bounds__(256) transform_rotate_quant_v2_kernel(const float* __restrict__ x, const float* __restrict__ smooth,
                                                                        SignMask sm, __half* __restrict__ out, __half* __restrict__ rotated,
                                                                        size_t n_chunks, int cpr, Modulate mod) {
    // [cpr][ROW]: smooth[c] * sigma[c % 128]; rows padded by 16 floats so that the two lane sets of a
    // quarter-warp (adjacent chunk columns) read from different banks
    extern __shared__ float s_mul[];
    constexpr int LPG = 4, NV = 8, ROW = 144;
    // `smooth` and the sign mask are parameters of the model, not products of the previous kernel: the
    // table is built before pdl_wait(), i.e. while the previous kernel is still draining
    pdl_launch_dependents();
    for (int c = threadIdx.x; c < cpr * 128; c += blockDim.x) {
        const int e = c & 127;
        const float sv = smooth != nullptr ? __ldg(smooth + c) : 1.0f;
        s_mul[(c >> 7) * ROW + e] = ((sm.w[e >> 5] >> (e & 31)) & 1u) ? sv : -sv;      // sign flips are exact: (x*s)*sigma == x*(s*sigma)
    }
    __syncthreads();
    pdl_wait();
    // (Measured and rejected, profiles/r1_kbench.txt: requesting the first chunk before the barrier, or the
    // next chunk before computing this one, costs 32 live registers -- 100 instead of 61, or spills at 64 --
    // and loses 10-25 % at every size.)
    const int lane = threadIdx.x & 31;
    const int lig = lane % LPG;
    const size_t warp_global = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const size_t n_warps = (size_t(gridDim.x) * blockDim.x) >> 5;
    const uint64_t neg1 = pk(-1.0f, -1.0f);
    const float delta = tie_delta_kernel(uint32_t(warp_global >> 33));
    const float cinv = inv_sqrt128();
    const uint64_t cinv2 = pk(cinv, cinv);
    const float sg1 = (lig & 1) ? -1.0f : 1.0f, sg2 = (lig & 2) ? -1.0f : 1.0f;

    for (size_t cbase = warp_global * 8; cbase < n_chunks; cbase += n_warps * 8) {
        const size_t c = cbase + lane / LPG;
        const bool valid = c < n_chunks;
        const size_t off = c * 128;
        const int ccol = valid ? int(c % size_t(cpr)) : 0;
        const float* mrow = s_mul + ccol * ROW;
        size_t moff = 0;
        if constexpr (MOD) moff = (valid ? (c / size_t(cpr)) / mod.rows_per_batch : 0) * (size_t(cpr) * 128) + size_t(ccol) * 128;
        uint64_t P[16];
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            uint4 u = make_uint4(0u, 0u, 0u, 0u);
            if (valid) u = ldg_stream(x + off + (j * LPG + lig) * 4);
            const float4 m4 = *reinterpret_cast<const float4*>(mrow + (j * LPG + lig) * 4);
            uint64_t xa = (uint64_t(u.y) << 32) | u.x, xb = (uint64_t(u.w) << 32) | u.z;
            if constexpr (MOD) {
                const size_t mo = moff + (j * LPG + lig) * 4;
                const float4 sc = __ldg(reinterpret_cast<const float4*>(mod.scale + mo));
                const float4 sh = __ldg(reinterpret_cast<const float4*>(mod.shift + mo));
                const uint64_t one2 = pk(mod.one, mod.one);
                // .mul(scale.add(1)).add_(shift).  The product uses the SCALAR mul.rn.f32 (never contracted):
                // ptxas fuses mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 -- even through a *1 -- which would skip
                // the rounding of the product that the reference's separate ATen kernels perform.
                const F2 a0 = unpk(fadd2(pk(sc.x, sc.y), one2)), a1 = unpk(fadd2(pk(sc.z, sc.w), one2));
                const F2 x0 = unpk(xa), x1 = unpk(xb);
                xa = fadd2(pk(__fmul_rn(x0.lo, a0.lo), __fmul_rn(x0.hi, a0.hi)), pk(sh.x, sh.y));
                xb = fadd2(pk(__fmul_rn(x1.lo, a1.lo), __fmul_rn(x1.hi, a1.hi)), pk(sh.z, sh.w));
            }
            // x * (s * sigma): basic_var.py:263 `.mul(s)` in fp32, then the sign row of Q
            P[2 * j] = fmul2(xa, pk(m4.x, m4.y));
            P[2 * j + 1] = fmul2(xb, pk(m4.z, m4.w));
        }
        // index bit 0: inside a packed pair
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const F2 f = unpk(P[i]);
            P[i] = pk(f.lo + f.hi, f.lo - f.hi);
        }
        // index bits 1, 4, 5, 6: between packed registers (distance 1, 2, 4, 8 in P[])
#pragma unroll
        for (int h = 1; h < 16; h <<= 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if ((i & h) == 0) {
                    const uint64_t a = P[i], b = P[i + h];
                    P[i] = fadd2(a, b);
                    P[i + h] = ffma2(b, neg1, a);
                }
            }
        }
        // index bits 2, 3: across the 4 lanes of the set
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
            const float sg = o == 1 ? sg1 : sg2;
            const uint64_t sgp = pk(sg, sg);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const F2 f = unpk(P[i]);
                const uint64_t q = pk(__shfl_xor_sync(0xffffffffu, f.lo, o), __shfl_xor_sync(0xffffffffu, f.hi, o));
                P[i] = ffma2(P[i], sgp, q);            // upper lane: partner - mine ; lower lane: mine + partner (exact: * +-1)
            }
        }
        // / fl32(sqrt(128)), rounded to fp16: the fp16 GEMM output of the reference
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = pack_h2_u64(fmul2(P[i], cinv2));
        if (valid && rotated != nullptr) {
#pragma unroll
            for (int j = 0; j < NV; ++j) stg_stream(rotated + off + (j * LPG + lig) * 4, make_uint2(w[2 * j], w[2 * j + 1]));
        }
        bool ok = true;
        float s = 0.0f;
        if constexpr (QUANT) ok = sym_quant_tile_h16<FMT, LPG, 16>(w, s, delta);
        if (valid) {
#pragma unroll
            for (int j = 0; j < NV; ++j) stg_stream(out + off + (j * LPG + lig) * 4, make_uint2(w[2 * j], w[2 * j + 1]));
            // irregular scale (zero / subnormal / inf / NaN): `out` now holds this lane's rotated values;
            // quantize them in place with the literal reference sequence
            if (!ok) literal_sym_h16(out + off, out + off, lig, LPG, 4, NV, s, SymFmt<FMT>::GT);
        }
    }
}

// Weight side: one warp per (row, chunk); lane l holds elements 4l..4l+3 in fp64.
__global__ void __launch_bounds__(256) transform_rotate_weight_kernel(const float* __restrict__ w, const float* __restrict__ smooth,
                                                                      SignMask sm, float* __restrict__ w_out, size_t n_chunks,
                                                                      int chunks_per_row) {
    const int lane = threadIdx.x & 31;
    const size_t warp_global = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const size_t n_warps = (size_t(gridDim.x) * blockDim.x) >> 5;
    for (size_t c = warp_global; c < n_chunks; c += n_warps) {
        const float4 w4 = *reinterpret_cast<const float4*>(w + c * 128 + 4 * lane);
        float f[4] = {w4.x, w4.y, w4.z, w4.w};
        if (smooth != nullptr) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(smooth + size_t(c % chunks_per_row) * 128 + 4 * lane));
            f[0] = __fdiv_rn(f[0], s4.x); f[1] = __fdiv_rn(f[1], s4.y);          // transform_model_utils.py:12 (fp32)
            f[2] = __fdiv_rn(f[2], s4.z); f[3] = __fdiv_rn(f[3], s4.w);
        }
        double d[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = 4 * lane + k;
            const bool plus = (sm.w[e >> 5] >> (e & 31)) & 1u;
            d[k] = plus ? double(f[k]) : -double(f[k]);
        }
        // strides 1, 2 in registers
        { const double a = d[0] + d[1], b = d[0] - d[1], c2 = d[2] + d[3], e2 = d[2] - d[3];
          d[0] = a + c2; d[2] = a - c2; d[1] = b + e2; d[3] = b - e2; }
        // strides 4..64 across the 32 lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const bool upper = lane & o;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double p = __shfl_xor_sync(0xffffffffu, d[k], o);
                d[k] = upper ? p - d[k] : d[k] + p;
            }
        }
        const double c64 = double(11.313708305358887f);       // fl32(sqrt(128)) widened, hadamard_utils.py:85
        float4 o4;
        o4.x = float(d[0] / c64); o4.y = float(d[1] / c64); o4.z = float(d[2] / c64); o4.w = float(d[3] / c64);
        *reinterpret_cast<float4*>(w_out + c * 128 + 4 * lane) = o4;
    }
}

}  // namespace fpq

using namespace fpq;

static int launch_rotate_quant(const float* x, const Modulate* mod, const float* smooth, const uint32_t* sign_bits_host, void* out, void* rotated,
                               size_t n_rows, size_t n_cols, int format, void* stream) {
    if (n_cols == 0 || n_cols % 128 != 0 || !sign_bits_host || (n_rows && (!x || !out))) return FPQ_ERR_ARG;
    if (format < -1 || format >= FPQ_NUM_SYM_FORMATS) return FPQ_ERR_ARG;
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(smooth)) & 15) || (reinterpret_cast<uintptr_t>(out) & 7) ||
        (reinterpret_cast<uintptr_t>(rotated) & 7))
        return FPQ_ERR_ARG;
    if (mod != nullptr) {
        if (!mod->scale || !mod->shift || mod->rows_per_batch == 0 || n_rows % mod->rows_per_batch != 0) return FPQ_ERR_ARG;
        if ((reinterpret_cast<uintptr_t>(mod->scale) | reinterpret_cast<uintptr_t>(mod->shift)) & 15) return FPQ_ERR_ARG;
    }
    if (n_rows == 0) return FPQ_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SignMask sm;
    for (int i = 0; i < 4; ++i) sm.w[i] = sign_bits_host[i];
    const int cpr = int(n_cols / 128);
    __half* o = static_cast<__half*>(out);
    __half* rot = static_cast<__half*>(rotated);
#if FPQ_ROT_V2
    // Small launches (the first four stages of a VAR pass) are latency-bound: the first layout has no
    // shared-memory table to build and twice as many lane sets per chunk, and is 5-35 % faster there
    // (profiles/r1_stagebench.txt); from ~25 k chunks on the second layout wins (fewer instructions).
    const bool small = n_rows * size_t(cpr) <= 24576;
    // FPQ_ROT_MOD_V2=1 keeps the second layout for the modulate variant too (it re-reads scale/shift per chunk through L1:
    // 3.6 TB/s-equivalent against the register-table variant of the first layout, see profiles/r1_kbench.txt)
    static const bool mod_v2 = getenv("FPQ_ROT_MOD_V2") != nullptr;
    if (!small && (mod == nullptr || mod_v2) && size_t(cpr) * 144 * sizeof(float) <= 48 * 1024) {
        const size_t n_chunks = n_rows * size_t(cpr);
        const unsigned grid = grid_for(n_chunks, 64, 4);              // 8 warps x 8 chunks per block and trip
        const size_t smem = size_t(cpr) * 144 * sizeof(float);
        const Modulate m = mod ? *mod : Modulate{nullptr, nullptr, 1, 1.0f};
#define FPQ_TRQ2(F, Q)                                                                                                                     \
    if (mod) launch_pdl(transform_rotate_quant_v2_kernel<F, Q, true>, grid, 256, smem, st, x, smooth, sm, o, rot, n_chunks, cpr, m);       \
    else launch_pdl(transform_rotate_quant_v2_kernel<F, Q, false>, grid, 256, smem, st, x, smooth, sm, o, rot, n_chunks, cpr, m)
        switch (format) {
            case -1: FPQ_TRQ2(0, false); break;
            case FPQ_FMT_E2M1: FPQ_TRQ2(FPQ_FMT_E2M1, true); break;
            case FPQ_FMT_E1M2: FPQ_TRQ2(FPQ_FMT_E1M2, true); break;
            case FPQ_FMT_E3M0: FPQ_TRQ2(FPQ_FMT_E3M0, true); break;
            case FPQ_FMT_E2M3: FPQ_TRQ2(FPQ_FMT_E2M3, true); break;
            default: FPQ_TRQ2(FPQ_FMT_E3M2, true); break;
        }
#undef FPQ_TRQ2
        return finish_launch();
    }
#endif
    // first layout: lane sets, one per (chunk column, row phase); enough to fill every SM's thread slots (the
    // modulate variant holds three operand tables in registers: 2 resident CTAs instead of 4)
    const size_t max_sets = size_t(sm_count()) * (mod ? 1024 : 2048) / 8;
    size_t sets_per_col = max_sets / size_t(cpr);
    if (sets_per_col < 1) sets_per_col = 1;
    if (sets_per_col > n_rows) sets_per_col = n_rows;
    // equal work per set: with `trips` passes over the rows, use just enough sets that every pass is full
    // (25600 rows on 2525 sets would run 10 full passes and one at 10 %)
    const size_t trips = (n_rows + sets_per_col - 1) / sets_per_col;
    sets_per_col = (n_rows + trips - 1) / trips;
    const size_t n_sets = sets_per_col * size_t(cpr);
    const unsigned grid = unsigned((n_sets + 31) / 32);            // 32 lane sets per 256-thread block
    const Modulate m1 = mod ? *mod : Modulate{nullptr, nullptr, 1, 1.0f};
#define FPQ_TRQ(F, Q)                                                                                                                     \
    if (mod) launch_pdl(transform_rotate_quant_kernel<F, Q, true>, grid, 256, 0, st, x, smooth, sm, o, rot, n_rows, cpr, sets_per_col, m1); \
    else launch_pdl(transform_rotate_quant_kernel<F, Q, false>, grid, 256, 0, st, x, smooth, sm, o, rot, n_rows, cpr, sets_per_col, m1)
    switch (format) {
        case -1: FPQ_TRQ(0, false); break;
        case FPQ_FMT_E2M1: FPQ_TRQ(FPQ_FMT_E2M1, true); break;
        case FPQ_FMT_E1M2: FPQ_TRQ(FPQ_FMT_E1M2, true); break;
        case FPQ_FMT_E3M0: FPQ_TRQ(FPQ_FMT_E3M0, true); break;
        case FPQ_FMT_E2M3: FPQ_TRQ(FPQ_FMT_E2M3, true); break;
        default: FPQ_TRQ(FPQ_FMT_E3M2, true); break;
    }
#undef FPQ_TRQ
    return finish_launch();
}

extern "C" int fpq_transform_rotate_quant(const float* x, const float* smooth, const uint32_t* sign_bits_host, void* out, void* rotated,
                                          size_t n_rows, size_t n_cols, int format, void* stream) {
    return launch_rotate_quant(x, nullptr, smooth, sign_bits_host, out, rotated, n_rows, n_cols, format, stream);
}

extern "C" int fpq_modulate_transform_rotate_quant(const float* x, const float* scale, const float* shift, size_t rows_per_batch,
                                                   const float* smooth, const uint32_t* sign_bits_host, void* out, void* rotated,
                                                   size_t n_rows, size_t n_cols, int format, int flags, void* stream) {
    if (flags & ~FPQ_MOD_GAIN) return FPQ_ERR_ARG;
    const Modulate m{scale, shift, rows_per_batch, (flags & FPQ_MOD_GAIN) ? -0.0f : 1.0f};
    return launch_rotate_quant(x, &m, smooth, sign_bits_host, out, rotated, n_rows, n_cols, format, stream);
}

extern "C" int fpq_transform_rotate_weight(const float* w, const float* smooth, const uint32_t* sign_bits_host, float* w_out, size_t n_rows,
                                           size_t n_cols, void* stream) {
    if (n_cols == 0 || n_cols % 128 != 0 || !sign_bits_host || (n_rows && (!w || !w_out))) return FPQ_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(smooth) | reinterpret_cast<uintptr_t>(w_out)) & 15) return FPQ_ERR_ARG;
    if (n_rows == 0) return FPQ_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SignMask sm;
    for (int i = 0; i < 4; ++i) sm.w[i] = sign_bits_host[i];
    const int cpr = int(n_cols / 128);
    const size_t n_chunks = n_rows * size_t(cpr);
    transform_rotate_weight_kernel<<<grid_for(n_chunks, 8, 64), 256, 0, st>>>(w, smooth, sm, w_out, n_chunks, cpr);
    return finish_launch();
}
