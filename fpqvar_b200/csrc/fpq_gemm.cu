// Packed low-bit operands and the REAL low-bit GEMM on the 5th-generation tensor cores (not in the
// reference, which fake-quantizes and then calls F.linear on fp16 tensors: models_fp_quant_transform_rotate/quant_utils.py:764-769).
//
// Operand format ("codes"): the quantizer's grid value q of every element as ONE e4m3 byte -- every FP4 / FP6 grid of the
// reference (e2m1, e1m2, e3m0, e2m3, e3m2) is a subset of e4m3, so the byte holds q exactly -- plus the fp32 scale of every
// (row, 128-group).  The bytes lie in the order the tensor core reads them, so that a (128 rows x 128 K) operand tile is 16 KB
// of CONTIGUOUS memory that one 1-D bulk copy (cp.async.bulk, the TMA engine; no tensor map) lands in shared memory as the
// canonical K-major no-swizzle UMMA layout:
//     codes [K/128 slabs][rows_pad/8 row blocks][8 chunks of 16 K][8 rows][16 bytes]        rows_pad = rows rounded up to 128
//     scales[K/128 slabs][rows_pad]                                                        fp32, 0 for the padding rows
// At rest the FP4 formats shrink to 4 bits per element (fpq_codes_to_nibbles / fpq_nibbles_to_codes), same order.
//
// GEMM: C[m, n] = sum over slabs t of  sa[t, m] * sw[t, n] * P_t[m, n],   P_t = sum over the slab's 128 k of qa * qw,
// P_t by tcgen05.mma kind::f8f6f4 (e4m3 x e4m3, fp32 accumulate in tensor memory; exact: every product is a multiple of 2^-8
// and the sums stay far below 2^24 of them), the scale product by the epilogue warps on the CUDA cores while the tensor core
// works on the next slab (two accumulators in tensor memory).  The fp32 operation order is fixed -- acc = fma(P*sa, sw, acc),
// slabs ascending -- so that oracle/gemm_codes.c reproduces C bit for bit.
//
// One CTA per 128 x 128 tile of C, six warps: 0 = bulk-copy producer, 1 = MMA issuer (one lane), 2..5 = epilogue (each owns the
// 32 tensor-memory lanes of its warp-id quarter).  Ring of shared-memory stages {A tile, B tile, sa, sw} with full / empty
// mbarriers; tcgen05.commit hands a stage back and publishes an accumulator.
#include "fpq_common.cuh"
#include "fpq_stream.cuh"
#include <cuda_fp8.h>

namespace fpq {

namespace {

constexpr int GK = 128;                      // K slab: the reference's quantization group
constexpr int TM = 128, TN = 128;            // C tile
constexpr uint32_t BLK_BYTES = 1024;         // one (8 rows x 128 K) block of codes
constexpr uint32_t A_BYTES = TM * GK, B_BYTES = TN * GK, SA_BYTES = TM * 4, SW_BYTES = TN * 4;
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES + SA_BYTES + SW_BYTES;       // 33 792 = 33 * 1024
constexpr int MAX_STAGES = 6;
constexpr int GEMM_THREADS = 192;
constexpr uint32_t TMEM_COLS = 256;          // two 128-column fp32 accumulators

// ---------------------------------------------------------------------------------------------------------------------
// quantizer -> codes.  One warp per (8 rows x 128 K) block: lane = (row & 7) + 8 * (chunk & 3), two chunks of 16 elements per
// lane, so that a warp reads 8 x 128 contiguous bytes (fp16) per instruction pair and writes 512 contiguous bytes per store.
// Arithmetic: the reference's own sequence (qu.py:313-330 and Appendix A of SURVEY.md), with true divisions.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T> struct Chunk16;
template <> struct Chunk16<__half> {
    static __device__ __forceinline__ void load(const __half* p, float (&v)[16]) {
        const uint4 a = ldg_stream(p), b = ldg_stream(p + 8);
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[2 * i] = h2f(uint16_t(w[i] & 0xffffu));
            v[2 * i + 1] = h2f(uint16_t(w[i] >> 16));
        }
    }
    static __device__ __forceinline__ float scale(float amax, float vmax) { return __half2float(__float2half_rn(__fdiv_rn(amax, vmax))); }
    static __device__ __forceinline__ float norm(float x, float s) { return __half2float(__float2half_rn(__fdiv_rn(x, s))); }
};
template <> struct Chunk16<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[16]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 a = ldg_stream(p + 4 * j);
            v[4 * j] = __uint_as_float(a.x); v[4 * j + 1] = __uint_as_float(a.y);
            v[4 * j + 2] = __uint_as_float(a.z); v[4 * j + 3] = __uint_as_float(a.w);
        }
    }
    static __device__ __forceinline__ float scale(float amax, float vmax) { return __fdiv_rn(amax, vmax); }
    static __device__ __forceinline__ float norm(float x, float s) { return __fdiv_rn(x, s); }
};

__device__ __forceinline__ uint32_t e4m3x4(float a, float b, float c, float d) {
    const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
    const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
    return lo | (hi << 16);
}

template <typename T, class HG>
__global__ void __launch_bounds__(256) pack_codes_kernel(const T* __restrict__ x, size_t rows, size_t rows_pad, size_t k,
                                                         uint8_t* __restrict__ codes, float* __restrict__ scales) {
    const int lane = threadIdx.x & 31;
    const int r = lane & 7, kq = lane >> 3;
    const size_t slabs = k / GK, row_blocks = rows_pad / 8;
    const size_t n_tasks = slabs * row_blocks;
    const size_t warp0 = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, n_warps = (size_t(gridDim.x) * blockDim.x) >> 5;
    for (size_t task = warp0; task < n_tasks; task += n_warps) {
        const size_t rb = task / slabs, slab = task % slabs;         // consecutive warps: consecutive slabs of the same rows
        const size_t row = rb * 8 + r;
        const bool live = row < rows;
        float v[2][16];
        float amax = 0.0f;
        bool nan = false;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (live) Chunk16<T>::load(x + row * k + slab * GK + size_t(kq + 4 * j) * 16, v[j]);
            else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[j][i] = 0.0f;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                amax = fmaxf(amax, fabsf(v[j][i]));
                nan |= v[j][i] != v[j][i];
            }
        }
        // the row's group lives in the lanes with the same (lane & 7)
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 8));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 16));
        unsigned nanmask = __ballot_sync(0xffffffffu, nan);
        nanmask |= nanmask >> 16;
        nanmask |= nanmask >> 8;
        if ((nanmask >> r) & 1u) amax = __int_as_float(0x7fc00000);          // torch's max propagates NaN
        const float s = Chunk16<T>::scale(amax, HG::VMAX);
        uint8_t* blk = codes + (slab * row_blocks + rb) * BLK_BYTES;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float q[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) q[i] = round_any_sym<HG, TIE_KERNEL>(Chunk16<T>::norm(v[j][i], s));
            uint4 o;
            o.x = e4m3x4(q[0], q[1], q[2], q[3]);
            o.y = e4m3x4(q[4], q[5], q[6], q[7]);
            o.z = e4m3x4(q[8], q[9], q[10], q[11]);
            o.w = e4m3x4(q[12], q[13], q[14], q[15]);
            *reinterpret_cast<uint4*>(blk + (kq + 4 * j) * 128 + r * 16) = o;
        }
        if (kq == 0) scales[slab * rows_pad + row] = live ? s : 0.0f;
    }
}

// codes -> values: out[row, c] = T( fl32(q) * scale ), the reference's last step (qu.py:328-329).  One thread per 16-byte chunk.
template <typename T>
__global__ void __launch_bounds__(256) unpack_codes_kernel(const uint8_t* __restrict__ codes, const float* __restrict__ scales, size_t rows,
                                                           size_t rows_pad, size_t k, T* __restrict__ out) {
    const size_t slabs = k / GK, row_blocks = rows_pad / 8;
    const size_t n_chunks = slabs * row_blocks * 64;
    for (size_t c = size_t(blockIdx.x) * blockDim.x + threadIdx.x; c < n_chunks; c += size_t(gridDim.x) * blockDim.x) {
        const size_t blk = c >> 6;
        const int kc = int(c >> 3) & 7, r = int(c) & 7;
        const size_t slab = blk / row_blocks, rb = blk % row_blocks;
        const size_t row = rb * 8 + r;
        if (row >= rows) continue;
        const uint4 u = *reinterpret_cast<const uint4*>(codes + c * 16);
        const float s = scales[slab * rows_pad + row];
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
        T* o = out + row * k + slab * GK + kc * 16;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int b = 0; b < 4; b += 2) {
                const __half2_raw hr = __nv_cvt_fp8x2_to_halfraw2(__nv_fp8x2_storage_t((w[i] >> (8 * b)) & 0xffffu), __NV_E4M3);
                const float q0 = __half2float(__ushort_as_half(hr.x)), q1 = __half2float(__ushort_as_half(hr.y));
                o[4 * i + b] = T(__fmul_rn(q0, s));
                o[4 * i + b + 1] = T(__fmul_rn(q1, s));
            }
        }
    }
}

// 4-bit storage of the FP4 formats: nibble = sign << 3 | index of |q| in the format's non-negative half grid (ascending);
// byte i holds elements 2i (low nibble) and 2i + 1 of the codes array, whatever its order.
struct NibbleLut { uint8_t e4m3[16]; };

__global__ void __launch_bounds__(256) codes_to_nibbles_kernel(const uint8_t* __restrict__ codes, size_t n_words, NibbleLut lut,
                                                               uint8_t* __restrict__ nib) {
    for (size_t w = size_t(blockIdx.x) * blockDim.x + threadIdx.x; w < n_words; w += size_t(gridDim.x) * blockDim.x) {
        const uint2 u = reinterpret_cast<const uint2*>(codes)[w];
        const uint32_t in[2] = {u.x, u.y};
        uint32_t out = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint8_t b = uint8_t(in[i >> 2] >> (8 * (i & 3)));
            uint32_t idx = 0;
#pragma unroll
            for (int j = 1; j < 8; ++j) idx = (b & 0x7f) == lut.e4m3[j] ? j : idx;
            out |= (idx | ((b >> 7) << 3)) << (4 * i);
        }
        reinterpret_cast<uint32_t*>(nib)[w] = out;
    }
}

__global__ void __launch_bounds__(256) nibbles_to_codes_kernel(const uint8_t* __restrict__ nib, size_t n_words, NibbleLut lut,
                                                               uint8_t* __restrict__ codes) {
    for (size_t w = size_t(blockIdx.x) * blockDim.x + threadIdx.x; w < n_words; w += size_t(gridDim.x) * blockDim.x) {
        const uint32_t in = reinterpret_cast<const uint32_t*>(nib)[w];
        uint32_t out[2] = {0, 0};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t n = (in >> (4 * i)) & 0xfu;
            out[i >> 2] |= uint32_t(lut.e4m3[n]) << (8 * (i & 3));
        }
        reinterpret_cast<uint2*>(codes)[w] = make_uint2(out[0], out[1]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// tcgen05 primitives
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, 128 x 128 x 32, e4m3 operands, fp32 accumulate
__device__ __forceinline__ void tc_mma_f8(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (base lane + t)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no swizzle: core matrix = 8 rows x 16 bytes (128 contiguous bytes); `lbo` = byte step between the two core
// matrices one MMA reads along K, `sbo` = byte step between 8-row groups; descriptor version 1 (Blackwell), layout type 0.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return uint64_t((smem_addr >> 4) & 0x3fffu) | (uint64_t((lbo >> 4) & 0x3fffu) << 16) | (uint64_t((sbo >> 4) & 0x3fffu) << 32) |
           (uint64_t(1) << 46);
}
// instruction descriptor: D fp32 (bits 4-5 = 1), A / B e4m3 (format 0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC_E4M3_128x128 = (1u << 4) | (uint32_t(TN >> 3) << 17) | (uint32_t(TM >> 4) << 24);

struct GemmArgs {
    const uint8_t* a_codes;
    const float* a_scales;
    const uint8_t* b_codes;
    const float* b_scales;
    const float* bias;       // fp32 [n] or null
    void* c;
    size_t ldc;              // elements
    size_t m, n;             // valid rows / columns of C
    size_t m_pad, n_pad;     // padded row counts of the two code arrays (multiples of 128)
    uint32_t slabs;          // K / 128
    uint32_t stages;
    uint32_t lbo, sbo;
};

template <typename OutT>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_codes_kernel(const GemmArgs g) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;                 // stage bases on 1 KB
    uint8_t* const stage0 = smem_raw + (smem0 - smem_u32(smem_raw));
    const uint32_t stages = g.stages, slabs = g.slabs;

    const size_t tiles_n = g.n_pad / TN;
    const size_t tm = blockIdx.x / tiles_n, tn = blockIdx.x % tiles_n;            // consecutive CTAs share the A panel

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < stages; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1 + 4);          // the MMA's commit + one arrive per epilogue warp (scales read)
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bar_tfull[a], 1);
            mbar_init(&bar_tempty[a], 4);
        }
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===== producer: four bulk copies per slab =====
        if (lane == 0) {
            const size_t a_blocks = g.m_pad / 8, b_blocks = g.n_pad / 8;
            for (uint32_t i = 0; i < slabs; ++i) {
                const uint32_t s = i % stages, n = i / stages;
                if (n > 0) mbar_wait(&bar_empty[s], (n - 1) & 1);
                uint8_t* st = stage0 + size_t(s) * STAGE_BYTES;
                mbar_arrive_expect_tx(&bar_full[s], STAGE_BYTES);
                bulk_load(st, g.a_codes + (size_t(i) * a_blocks + tm * (TM / 8)) * BLK_BYTES, A_BYTES, &bar_full[s]);
                bulk_load(st + A_BYTES, g.b_codes + (size_t(i) * b_blocks + tn * (TN / 8)) * BLK_BYTES, B_BYTES, &bar_full[s]);
                bulk_load(st + A_BYTES + B_BYTES, g.a_scales + size_t(i) * g.m_pad + tm * TM, SA_BYTES, &bar_full[s]);
                bulk_load(st + A_BYTES + B_BYTES + SA_BYTES, g.b_scales + size_t(i) * g.n_pad + tn * TN, SW_BYTES, &bar_full[s]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer: four 128 x 128 x 32 MMAs per slab into accumulator (slab & 1) =====
        for (uint32_t i = 0; i < slabs; ++i) {
            const uint32_t s = i % stages, n = i / stages, a = i & 1, u = i >> 1;
            if (u > 0) mbar_wait(&bar_tempty[a], (u - 1) & 1);
            mbar_wait(&bar_full[s], n & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sa_addr = smem0 + s * STAGE_BYTES;
                const uint64_t da = umma_desc(sa_addr, g.lbo, g.sbo), db = umma_desc(sa_addr + A_BYTES, g.lbo, g.sbo);
#pragma unroll
                for (uint32_t kk = 0; kk < GK / 32; ++kk)      // 32 K = two core matrices = 256 bytes further on
                    tc_mma_f8(tmem_base + a * TN, da + kk * (256u >> 4), db + kk * (256u >> 4), IDESC_E4M3_128x128, kk);
                tc_commit(&bar_empty[s]);
                tc_commit(&bar_tfull[a]);
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: acc = fma(P * sa, sw, acc), one C row per thread =====
        const int quarter = warp & 3;                         // the tensor-memory lanes this warp may touch
        const int row_in_tile = quarter * 32 + lane;
        float acc[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[j] = 0.0f;
        for (uint32_t i = 0; i < slabs; ++i) {
            const uint32_t s = i % stages, n = i / stages, a = i & 1, u = i >> 1;
            const uint8_t* st = stage0 + size_t(s) * STAGE_BYTES;
            mbar_wait(&bar_full[s], n & 1);                   // the scales of this slab are in shared memory
            const float sa = reinterpret_cast<const float*>(st + A_BYTES + B_BYTES)[row_in_tile];
            const float4* sw4 = reinterpret_cast<const float4*>(st + A_BYTES + B_BYTES + SA_BYTES);
            mbar_wait(&bar_tfull[a], u & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < TN / 32; ++c) {
                uint32_t v[32];
                tc_ld32(tmem_base + (uint32_t(quarter * 32) << 16) + a * TN + c * 32, v);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 w = sw4[c * 8 + j / 4];
                    acc[c * 32 + j + 0] = __fmaf_rn(__fmul_rn(__uint_as_float(v[j + 0]), sa), w.x, acc[c * 32 + j + 0]);
                    acc[c * 32 + j + 1] = __fmaf_rn(__fmul_rn(__uint_as_float(v[j + 1]), sa), w.y, acc[c * 32 + j + 1]);
                    acc[c * 32 + j + 2] = __fmaf_rn(__fmul_rn(__uint_as_float(v[j + 2]), sa), w.z, acc[c * 32 + j + 2]);
                    acc[c * 32 + j + 3] = __fmaf_rn(__fmul_rn(__uint_as_float(v[j + 3]), sa), w.w, acc[c * 32 + j + 3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&bar_tempty[a]);
                mbar_arrive(&bar_empty[s]);
            }
        }
        const size_t row = tm * TM + row_in_tile, col0 = tn * TN;
        if (row < g.m) {
            OutT* crow = static_cast<OutT*>(g.c) + row * g.ldc + col0;
#pragma unroll
            for (int j = 0; j < TN; j += 8) {
                if (col0 + j >= g.n) break;                   // n is a multiple of 8 (checked by the launcher)
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = acc[j + e] + (g.bias ? __ldg(g.bias + col0 + j + e) : 0.0f);
                if constexpr (sizeof(OutT) == 2) {
                    *reinterpret_cast<uint4*>(crow + j) = make_uint4(pack_h2(o[0], o[1]), pack_h2(o[2], o[3]), pack_h2(o[4], o[5]), pack_h2(o[6], o[7]));
                } else {
                    *reinterpret_cast<float4*>(crow + j) = make_float4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<float4*>(crow + j + 4) = make_float4(o[4], o[5], o[6], o[7]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

bool half_grid_e4m3(int format, NibbleLut& lut) {
    // |q| of the three FP4 formats as e4m3 bytes (bias 7, 3 mantissa bits), ascending; entries 8..15 = the same with the sign bit
    static const uint8_t e2m1[8] = {0x00, 0x30, 0x38, 0x3c, 0x40, 0x44, 0x48, 0x4c};      // 0 .5 1 1.5 2 3 4 6
    static const uint8_t e1m2[8] = {0x00, 0x28, 0x30, 0x34, 0x38, 0x3a, 0x3c, 0x3e};      // 0 .25 .5 .75 1 1.25 1.5 1.75
    static const uint8_t e3m0[8] = {0x00, 0x28, 0x30, 0x38, 0x40, 0x48, 0x50, 0x58};      // 0 .25 .5 1 2 4 8 16
    const uint8_t* t = format == FPQ_FMT_E2M1 ? e2m1 : format == FPQ_FMT_E1M2 ? e1m2 : format == FPQ_FMT_E3M0 ? e3m0 : nullptr;
    if (!t) return false;
    for (int i = 0; i < 8; ++i) {
        lut.e4m3[i] = t[i];
        lut.e4m3[8 + i] = uint8_t(t[i] | 0x80);
    }
    lut.e4m3[8] = 0x00;           // the quantizer never produces -0 (the grids' zero is +0.0); decode nibble 8 as +0 too
    return true;
}

template <typename T>
int launch_pack(int format, const T* x, size_t rows, size_t rows_pad, size_t k, uint8_t* codes, float* scales, cudaStream_t st) {
    const size_t tasks = (k / GK) * (rows_pad / 8);
    const unsigned grid = grid_for(tasks, 8, 8);
    switch (format) {
        case FPQ_FMT_E2M1: pack_codes_kernel<T, HG_E2M1><<<grid, 256, 0, st>>>(x, rows, rows_pad, k, codes, scales); break;
        case FPQ_FMT_E1M2: pack_codes_kernel<T, HG_E1M2><<<grid, 256, 0, st>>>(x, rows, rows_pad, k, codes, scales); break;
        case FPQ_FMT_E3M0: pack_codes_kernel<T, HG_E3M0><<<grid, 256, 0, st>>>(x, rows, rows_pad, k, codes, scales); break;
        case FPQ_FMT_E2M3: pack_codes_kernel<T, HG_E2M3><<<grid, 256, 0, st>>>(x, rows, rows_pad, k, codes, scales); break;
        case FPQ_FMT_E3M2: pack_codes_kernel<T, HG_E3M2><<<grid, 256, 0, st>>>(x, rows, rows_pad, k, codes, scales); break;
        default: return FPQ_ERR_ARG;
    }
    return finish_launch();
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

}  // namespace fpq

using namespace fpq;

extern "C" size_t fpq_codes_rows_padded(size_t rows) { return (rows + TM - 1) / TM * TM; }

extern "C" int fpq_pack_codes(const void* x, size_t rows, size_t k, int in_dtype, int format, uint8_t* codes, float* scales, void* stream) {
    if (k == 0 || k % GK != 0 || (rows && (!x || !codes || !scales)) || !aligned16(x) || !aligned16(codes) || !aligned16(scales)) return FPQ_ERR_ARG;
    if (rows == 0) return FPQ_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t rows_pad = fpq_codes_rows_padded(rows);
    if (in_dtype == FPQ_F16) return launch_pack<__half>(format, static_cast<const __half*>(x), rows, rows_pad, k, codes, scales, st);
    if (in_dtype == FPQ_F32) return launch_pack<float>(format, static_cast<const float*>(x), rows, rows_pad, k, codes, scales, st);
    return FPQ_ERR_ARG;
}

extern "C" int fpq_unpack_codes(const uint8_t* codes, const float* scales, size_t rows, size_t k, int out_dtype, void* out, void* stream) {
    if (k == 0 || k % GK != 0 || (rows && (!out || !codes || !scales)) || !aligned16(codes)) return FPQ_ERR_ARG;
    if (rows == 0) return FPQ_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t rows_pad = fpq_codes_rows_padded(rows);
    const unsigned grid = grid_for(rows_pad * k / 16, 256, 8);
    if (out_dtype == FPQ_F16) unpack_codes_kernel<__half><<<grid, 256, 0, st>>>(codes, scales, rows, rows_pad, k, static_cast<__half*>(out));
    else if (out_dtype == FPQ_F32) unpack_codes_kernel<float><<<grid, 256, 0, st>>>(codes, scales, rows, rows_pad, k, static_cast<float*>(out));
    else return FPQ_ERR_ARG;
    return finish_launch();
}

extern "C" int fpq_codes_to_nibbles(const uint8_t* codes, size_t n_codes, int format, uint8_t* nibbles, void* stream) {
    NibbleLut lut;
    if (!half_grid_e4m3(format, lut)) return FPQ_ERR_UNSUPPORTED;
    if (n_codes % 8 != 0 || (n_codes && (!codes || !nibbles)) || !aligned16(codes) || !aligned16(nibbles)) return FPQ_ERR_ARG;
    if (n_codes == 0) return FPQ_OK;
    codes_to_nibbles_kernel<<<grid_for(n_codes / 8, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(codes, n_codes / 8, lut, nibbles);
    return finish_launch();
}

extern "C" int fpq_nibbles_to_codes(const uint8_t* nibbles, size_t n_codes, int format, uint8_t* codes, void* stream) {
    NibbleLut lut;
    if (!half_grid_e4m3(format, lut)) return FPQ_ERR_UNSUPPORTED;
    if (n_codes % 8 != 0 || (n_codes && (!codes || !nibbles)) || !aligned16(codes) || !aligned16(nibbles)) return FPQ_ERR_ARG;
    if (n_codes == 0) return FPQ_OK;
    nibbles_to_codes_kernel<<<grid_for(n_codes / 8, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(nibbles, n_codes / 8, lut, codes);
    return finish_launch();
}

extern "C" int fpq_gemm_codes(const uint8_t* a_codes, const float* a_scales, size_t m, const uint8_t* b_codes, const float* b_scales,
                              size_t n, size_t k, const float* bias, int out_dtype, void* c, size_t ldc, void* stream) {
    if (k == 0 || k % GK != 0 || k / GK > 0xffffffffull) return FPQ_ERR_ARG;
    if (m == 0 || n == 0) return FPQ_OK;
    if (!a_codes || !a_scales || !b_codes || !b_scales || !c) return FPQ_ERR_ARG;
    if (!aligned16(a_codes) || !aligned16(a_scales) || !aligned16(b_codes) || !aligned16(b_scales) || !aligned16(c)) return FPQ_ERR_ARG;
    if (n % 8 != 0 || ldc % 8 != 0 || ldc < n) return FPQ_ERR_ARG;
    if (out_dtype != FPQ_F16 && out_dtype != FPQ_F32) return FPQ_ERR_ARG;
    GemmArgs g;
    g.a_codes = a_codes; g.a_scales = a_scales; g.b_codes = b_codes; g.b_scales = b_scales; g.bias = bias;
    g.c = c; g.ldc = ldc; g.m = m; g.n = n;
    g.m_pad = fpq_codes_rows_padded(m); g.n_pad = fpq_codes_rows_padded(n);
    g.slabs = uint32_t(k / GK);
    g.stages = uint32_t(g_tun.gemm_stages);
    g.lbo = g_tun.gemm_desc_swap ? 1024u : 128u;
    g.sbo = g_tun.gemm_desc_swap ? 128u : 1024u;
    const size_t tiles = (g.m_pad / TM) * (g.n_pad / TN);
    if (tiles > 0x7fffffffull) return FPQ_ERR_UNSUPPORTED;
    const size_t smem = size_t(g.stages) * STAGE_BYTES + 1024;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static bool attr_set[64][2] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int oi = out_dtype == FPQ_F16 ? 0 : 1;
    if (!attr_set[dev & 63][oi]) {
        const size_t max_smem = size_t(MAX_STAGES) * STAGE_BYTES + 1024;
        cudaError_t e = oi == 0 ? cudaFuncSetAttribute(gemm_codes_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(max_smem))
                                : cudaFuncSetAttribute(gemm_codes_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(max_smem));
        if (e != cudaSuccess) return finish_launch();
        attr_set[dev & 63][oi] = true;
    }
    if (oi == 0) gemm_codes_kernel<__half><<<unsigned(tiles), GEMM_THREADS, smem, st>>>(g);
    else gemm_codes_kernel<float><<<unsigned(tiles), GEMM_THREADS, smem, st>>>(g);
    return finish_launch();
}
