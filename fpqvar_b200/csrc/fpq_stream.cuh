// TMA-staged streaming skeleton for the activation kernels.
//
// After the instruction diet of fpq_h16.cuh the kernels are latency-bound, not issue-bound (ncu r1b:
// long_scoreboard = 29-53 % of the stall samples at ~7 resident warps per scheduler).  Here the global
// loads leave the compute warps altogether: one producer lane per CTA streams contiguous tiles of
// groups into a ring of shared-memory stages with 1-D bulk async copies (cp.async.bulk -> UBLKCP, the
// TMA engine) that complete on an mbarrier; eight consumer warps wait on the "full" barrier, pull their
// groups into registers with conflict-free LDS.128, hand the stage back ("empty" barrier, one arrive
// per warp) and only then do the arithmetic and the streaming stores.  Up to STAGES-1 tiles per CTA are
// in flight regardless of what the consumers are doing.
#pragma once
#include "fpq_h16.cuh"

namespace fpq {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FPQ_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra FPQ_DONE;\n"
        "bra FPQ_WAIT;\n"
        "FPQ_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar`.  dst/src 16-byte aligned,
// bytes a multiple of 16.
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int ST_CONSUMER_WARPS = 8;
constexpr int ST_THREADS = 32 * (ST_CONSUMER_WARPS + 1);      // + one producer warp

// Ring state shared by producer and consumers.
template <int TILE_BYTES, int STAGES>
struct __align__(128) StreamSmem {
    static constexpr int NSTAGES = STAGES;
    unsigned char tile[STAGES][TILE_BYTES];
    uint64_t full[STAGES];
    uint64_t empty[STAGES];
};

// Producer loop: tiles t = blockIdx.x, += gridDim.x.  `tile_src(t)` / `tile_bytes(t)` describe tile t.
template <class Smem, class SrcFn, class BytesFn>
__device__ __forceinline__ void stream_producer(Smem& sm, size_t n_tiles, SrcFn tile_src, BytesFn tile_bytes) {
    uint32_t k = 0;
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++k) {
        const uint32_t s = k % Smem::NSTAGES, ph = (k / Smem::NSTAGES) & 1u;
        mbar_wait(&sm.empty[s], ph ^ 1u);                      // stage free (passes at once on the first lap)
        const uint32_t bytes = tile_bytes(t);
        mbar_arrive_expect_tx(&sm.full[s], bytes);
        bulk_load(sm.tile[s], tile_src(t), bytes, &sm.full[s]);
    }
}

template <class Smem>
__device__ __forceinline__ void stream_init(Smem& sm) {
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < Smem::NSTAGES; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], ST_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();
}

}  // namespace fpq
