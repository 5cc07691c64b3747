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
// product of a 2-point butterfly over every index bit, in any order.
#include "fpq_common.cuh"

namespace fpq {

struct SignMask { uint32_t w[4]; };    // bit e of the 128-bit mask set  <=>  sigma[e] = +1

// butterflies over the two register-resident index pairs (k: strides 1,2 ; j: strides 4,8 of v[])
template <typename T>
__device__ __forceinline__ void fwht16_regs(T (&v)[16]) {
#pragma unroll
    for (int h = 1; h < 16; h <<= 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if ((i & h) == 0) {
                const T a = v[i], b = v[i + h];
                v[i] = a + b;
                v[i + h] = a - b;
            }
        }
    }
}

// butterflies over the three lane bits of an 8-lane group
__device__ __forceinline__ void fwht_lanes8(float (&v)[16], int lig) {
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        const float sg = (lig & o) ? -1.0f : 1.0f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float p = __shfl_xor_sync(0xffffffffu, v[i], o);
            v[i] = fmaf(v[i], sg, p);          // upper lane: p - v ; lower lane: v + p  (exact: *+-1)
        }
    }
}

// 1 / fl32(sqrt(128)), rounded to fp32 (0x3DB504F3)
__device__ __forceinline__ float inv_sqrt128() { return __uint_as_float(0x3DB504F3u); }

template <int FMT, bool QUANT>
__global__ void __launch_bounds__(256) transform_rotate_quant_kernel(const float* __restrict__ x, const float* __restrict__ smooth,
                                                                     SignMask sm, __half* __restrict__ out, __half* __restrict__ rotated,
                                                                     size_t n_chunks, int chunks_per_row) {
    constexpr int LPG = 8;
    const int lane = threadIdx.x & 31;
    const int lig = lane % LPG;
    const size_t warp_global = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const size_t n_warps = (size_t(gridDim.x) * blockDim.x) >> 5;

    // sign of the 16 chunk positions this lane owns, as an xor mask on the fp32 sign bit
    uint32_t sgn[16];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) sgn[4 * j + k] = ((sm.w[j] >> (4 * lig + k)) & 1u) ? 0u : 0x80000000u;

    for (size_t cbase = warp_global * 4; cbase < n_chunks; cbase += n_warps * 4) {
        const size_t c = cbase + lane / LPG;
        const bool valid = c < n_chunks;
        float v[16];
        if (valid) {
            Vec16<float>::load(x + c * 128, lig, LPG, v);
            if (smooth != nullptr) {
                const float* sp = smooth + size_t(c % chunks_per_row) * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 s4 = __ldg(reinterpret_cast<const float4*>(sp + (j * LPG + lig) * 4));
                    v[4 * j + 0] = __fmul_rn(v[4 * j + 0], s4.x);     // basic_var.py:263 `.mul(s)`, fp32
                    v[4 * j + 1] = __fmul_rn(v[4 * j + 1], s4.y);
                    v[4 * j + 2] = __fmul_rn(v[4 * j + 2], s4.z);
                    v[4 * j + 3] = __fmul_rn(v[4 * j + 3], s4.w);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.0f;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(__float_as_uint(v[i]) ^ sgn[i]);
        fwht16_regs(v);
        fwht_lanes8(v, lig);
        const float cinv = inv_sqrt128();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __half2float(__float2half_rn(v[i] * cinv));   // the fp16 GEMM output of the reference
        if (valid && rotated != nullptr) store16<__half, 4>(rotated + c * 128, lig, LPG, v);
        if constexpr (QUANT) sym_quant_tile<__half, FMT, TIE_KERNEL, LPG>(v);
        if (valid) store16<__half, 4>(out + c * 128, lig, LPG, v);
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

extern "C" int fpq_transform_rotate_quant(const float* x, const float* smooth, const uint32_t* sign_bits_host, void* out, void* rotated,
                                          size_t n_rows, size_t n_cols, int format, void* stream) {
    if (n_cols == 0 || n_cols % 128 != 0 || !sign_bits_host || (n_rows && (!x || !out))) return FPQ_ERR_ARG;
    if (format < -1 || format >= FPQ_NUM_SYM_FORMATS) return FPQ_ERR_ARG;
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(smooth)) & 15) || (reinterpret_cast<uintptr_t>(out) & 7) ||
        (reinterpret_cast<uintptr_t>(rotated) & 7))
        return FPQ_ERR_ARG;
    if (n_rows == 0) return FPQ_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SignMask sm;
    for (int i = 0; i < 4; ++i) sm.w[i] = sign_bits_host[i];
    const int cpr = int(n_cols / 128);
    const size_t n_chunks = n_rows * size_t(cpr);
    const unsigned grid = grid_for(n_chunks, 32, 64);
    __half* o = static_cast<__half*>(out);
    __half* rot = static_cast<__half*>(rotated);
    switch (format) {
        case -1: transform_rotate_quant_kernel<0, false><<<grid, 256, 0, st>>>(x, smooth, sm, o, rot, n_chunks, cpr); break;
        case FPQ_FMT_E2M1: transform_rotate_quant_kernel<FPQ_FMT_E2M1, true><<<grid, 256, 0, st>>>(x, smooth, sm, o, rot, n_chunks, cpr); break;
        case FPQ_FMT_E1M2: transform_rotate_quant_kernel<FPQ_FMT_E1M2, true><<<grid, 256, 0, st>>>(x, smooth, sm, o, rot, n_chunks, cpr); break;
        case FPQ_FMT_E3M0: transform_rotate_quant_kernel<FPQ_FMT_E3M0, true><<<grid, 256, 0, st>>>(x, smooth, sm, o, rot, n_chunks, cpr); break;
        case FPQ_FMT_E2M3: transform_rotate_quant_kernel<FPQ_FMT_E2M3, true><<<grid, 256, 0, st>>>(x, smooth, sm, o, rot, n_chunks, cpr); break;
        default: transform_rotate_quant_kernel<FPQ_FMT_E3M2, true><<<grid, 256, 0, st>>>(x, smooth, sm, o, rot, n_chunks, cpr); break;
    }
    return finish_launch();
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
