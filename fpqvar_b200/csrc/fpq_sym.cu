// libfpq_b200 -- symmetric fake-quant kernels (reference rows a2, a4, a5, a6 of SURVEY.md section 8)
// Part of the C ABI of include/fpq_b200.h; no torch types here.
#include "fpq_common.cuh"

namespace fpq {

// ------------------------------------------------------------------------------------------
// Symmetric per-group fake quant, GS = 16*LPG elements per group
// ------------------------------------------------------------------------------------------
template <typename InT, typename OutT, int FMT, int TIE, int LPG>
__global__ void __launch_bounds__(256) fake_quant_group_kernel(const InT* __restrict__ x, OutT* __restrict__ out,
                                                               size_t n_groups, int clamp3) {
    constexpr int GS = 16 * LPG;
    constexpr int IN_VEC = 16 / sizeof(InT);
    constexpr int GROUPS_PER_WARP = 32 / LPG;
    const int lane = threadIdx.x & 31;
    const int lig = lane % LPG;                    // lane in group
    const size_t warp_global = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const size_t n_warps = (size_t(gridDim.x) * blockDim.x) >> 5;

    for (size_t gbase = warp_global * GROUPS_PER_WARP; gbase < n_groups; gbase += n_warps * GROUPS_PER_WARP) {
        const size_t g = gbase + lane / LPG;
        const bool valid = g < n_groups;
        float v[16];
        if (valid) {
            Vec16<InT>::load(x + g * GS, lig, LPG, v);
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.0f;
        }
        if (clamp3) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = clamp3_keep_nan(v[i]);
        }
        sym_quant_tile<InT, FMT, TIE, LPG>(v);
        if (valid) store16<OutT, IN_VEC>(out + g * GS, lig, LPG, v);
    }
}

// ------------------------------------------------------------------------------------------
// Symmetric per-row fake quant (per_token / per_channel), any row length.  One CTA per row
// (grid-stride over rows); pass 1 reduces the absmax, pass 2 re-reads the row (L2-resident,
// it was just streamed) and quantizes.  Scalar accesses unless the row is 16-byte aligned.
// ------------------------------------------------------------------------------------------
template <typename InT, typename OutT, int FMT, int TIE>
__global__ void __launch_bounds__(256) fake_quant_row_kernel(const InT* __restrict__ x, OutT* __restrict__ out,
                                                             size_t n_rows, size_t row_len, int clamp3) {
    using HG = typename SymFmt<FMT>::HG;
    __shared__ float red[32];
    constexpr int VEC = 16 / sizeof(InT);
    for (size_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
        const InT* xr = x + row * row_len;
        OutT* orow = out + row * row_len;
        const bool vec_ok = (row_len % VEC == 0) && ((reinterpret_cast<uintptr_t>(xr) & 15) == 0);
        float a = 0.0f;
        if (vec_ok) {
            for (size_t i = threadIdx.x; i < row_len / VEC; i += blockDim.x) {
                const uint4 u = *reinterpret_cast<const uint4*>(xr + i * VEC);
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if constexpr (sizeof(InT) == 2) {
                        float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
                        if (clamp3) { f.x = clamp3_keep_nan(f.x); f.y = clamp3_keep_nan(f.y); }
                        a = fmax_nan(a, fmax_nan(fabsf(f.x), fabsf(f.y)));
                    } else {
                        float f = __uint_as_float(w[k]);
                        if (clamp3) f = clamp3_keep_nan(f);
                        a = fmax_nan(a, fabsf(f));
                    }
                }
            }
        } else {
            for (size_t i = threadIdx.x; i < row_len; i += blockDim.x) {
                float f = load_elem(xr + i);
                if (clamp3) f = clamp3_keep_nan(f);
                a = fmax_nan(a, fabsf(f));
            }
        }
        a = block_max_nan(a, red);
        const float s = rnd_in<InT>(__fdiv_rn(a, HG::VMAX));
        const bool regular = scale_regular<InT>(s);
        const float r = regular ? __frcp_rn(s) : 0.0f;
        const GridTable& gt = c_grids[SymFmt<FMT>::GT];
        for (size_t i = threadIdx.x; i < row_len; i += blockDim.x) {
            float f = load_elem(xr + i);
            if (clamp3) f = clamp3_keep_nan(f);
            const float q = regular ? quant_elem_fast<InT, HG, TIE>(f, s, r) : quant_elem_literal<InT, TIE>(f, s, gt);
            store_elem(orow + i, q * s);
        }
    }
}

template <typename InT, typename OutT, int FMT, int TIE>
static int launch_sym(const void* x, void* out, size_t n_rows, size_t row_len, int clamp3, cudaStream_t st) {
    const InT* xi = static_cast<const InT*>(x);
    OutT* oo = static_cast<OutT*>(out);
    const bool pow2_group = (row_len == 128 || row_len == 64) &&
                            ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if constexpr (sizeof(InT) == 2 && sizeof(OutT) == 2 && TIE == TIE_KERNEL) {
        if (pow2_group && row_len == 128 && !clamp3) return launch_sym_h16(FMT, x, out, n_rows, st);
    }
    if (pow2_group) {
        const int lpg = int(row_len / 16);
        const size_t groups_per_block = (256 / 32) * (32 / lpg);
        const unsigned grid = grid_for(n_rows, groups_per_block, 64);
        if (lpg == 8) fake_quant_group_kernel<InT, OutT, FMT, TIE, 8><<<grid, 256, 0, st>>>(xi, oo, n_rows, clamp3);
        else fake_quant_group_kernel<InT, OutT, FMT, TIE, 4><<<grid, 256, 0, st>>>(xi, oo, n_rows, clamp3);
    } else {
        const unsigned grid = grid_for(n_rows, 1, 16);
        fake_quant_row_kernel<InT, OutT, FMT, TIE><<<grid, 256, 0, st>>>(xi, oo, n_rows, row_len, clamp3);
    }
    return finish_launch();
}

template <typename InT, typename OutT, int TIE>
static int dispatch_sym_fmt(int format, const void* x, void* out, size_t n_rows, size_t row_len, int clamp3, cudaStream_t st) {
    switch (format) {
        case FPQ_FMT_E2M1: return launch_sym<InT, OutT, FPQ_FMT_E2M1, TIE>(x, out, n_rows, row_len, clamp3, st);
        case FPQ_FMT_E1M2: return launch_sym<InT, OutT, FPQ_FMT_E1M2, TIE>(x, out, n_rows, row_len, clamp3, st);
        case FPQ_FMT_E3M0: return launch_sym<InT, OutT, FPQ_FMT_E3M0, TIE>(x, out, n_rows, row_len, clamp3, st);
        case FPQ_FMT_E2M3: return launch_sym<InT, OutT, FPQ_FMT_E2M3, TIE>(x, out, n_rows, row_len, clamp3, st);
        case FPQ_FMT_E3M2: return launch_sym<InT, OutT, FPQ_FMT_E3M2, TIE>(x, out, n_rows, row_len, clamp3, st);
        default: return FPQ_ERR_ARG;
    }
}

template <int TIE>
static int dispatch_sym_types(int in_dtype, int out_dtype, int format, const void* x, void* out, size_t n_rows, size_t row_len,
                              int clamp3, cudaStream_t st) {
    if (in_dtype == FPQ_F32 && out_dtype == FPQ_F32) return dispatch_sym_fmt<float, float, TIE>(format, x, out, n_rows, row_len, clamp3, st);
    if (in_dtype == FPQ_F32 && out_dtype == FPQ_F16) return dispatch_sym_fmt<float, __half, TIE>(format, x, out, n_rows, row_len, clamp3, st);
    if (in_dtype == FPQ_F16 && out_dtype == FPQ_F16) return dispatch_sym_fmt<__half, __half, TIE>(format, x, out, n_rows, row_len, clamp3, st);
    if (in_dtype == FPQ_F16 && out_dtype == FPQ_F32) return dispatch_sym_fmt<__half, float, TIE>(format, x, out, n_rows, row_len, clamp3, st);
    return FPQ_ERR_ARG;
}

}  // namespace fpq

using namespace fpq;

extern "C" int fpq_fake_quant(const void* x, void* out, size_t n_rows, size_t row_len, int in_dtype, int out_dtype, int format,
                              int tie_mode, unsigned flags, void* stream) {
    if (row_len == 0 || (n_rows && (!x || !out || x == out))) return FPQ_ERR_ARG;
    if (flags & ~FPQ_FLAG_CLAMP3) return FPQ_ERR_ARG;
    if (n_rows == 0) return FPQ_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int clamp3 = (flags & FPQ_FLAG_CLAMP3) ? 1 : 0;
    if (tie_mode == FPQ_TIE_KERNEL) return dispatch_sym_types<TIE_KERNEL>(in_dtype, out_dtype, format, x, out, n_rows, row_len, clamp3, st);
    if (tie_mode == FPQ_TIE_ARGMIN) return dispatch_sym_types<TIE_ARGMIN>(in_dtype, out_dtype, format, x, out, n_rows, row_len, clamp3, st);
    return FPQ_ERR_ARG;
}
