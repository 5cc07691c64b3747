// Clocks per tcgen05.mma (128 x 256 x 32, e4m3, operands in shared memory) by operand layout: does the K-major no-swizzle layout of
// fpq_gemm.cu (core matrices 128 B, 8-row groups 1024 B apart) feed the tensor core as fast as denser strides or the 128-byte swizzle?
// (development aid; build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fpqvar_b200/variants/umma_bench tools/umma_bench.cu)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void __launch_bounds__(128) bench(int iters, uint32_t lbo, uint32_t sbo, uint32_t layout_type, uint32_t kstep_bytes, uint32_t n_cols,
                                             unsigned long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x38383838u;   // e4m3 1.0
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t a_addr = (smem_u32(smem) + 1023u) & ~1023u, b_addr = a_addr + 32 * 1024;
        auto desc = [&](uint32_t addr) {
            return uint64_t((addr >> 4) & 0x3fffu) | (uint64_t((lbo >> 4) & 0x3fffu) << 16) | (uint64_t((sbo >> 4) & 0x3fffu) << 32) | (uint64_t(1) << 46) |
                   (uint64_t(layout_type) << 61);
        };
        const uint32_t idesc = (1u << 4) | ((n_cols >> 3) << 17) | (uint32_t(128 >> 4) << 24);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (uint32_t kk = 0; kk < 4; ++kk) {
                const uint64_t da = desc(a_addr + kk * kstep_bytes), db = desc(b_addr + kk * kstep_bytes);
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem + (i & 1) * 256),
                    "l"(da), "l"(db), "r"(idesc), "r"(kk)
                    : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static void run(const char* name, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t kstep, uint32_t n, unsigned long long* d) {
    const int iters = 2000;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; ++rep) bench<<<148, 128, 100 * 1024>>>(iters, lbo, sbo, layout, kstep, n, d);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    unsigned long long mx = 0;
    for (auto c : h) mx = c > mx ? c : mx;
    printf("%-70s N=%3u: %7.1f clk per MMA  (%s)\n", name, n, double(mx) / (iters * 4.0), cudaGetErrorString(e));
}

int main() {
    unsigned long long* d;
    cudaMalloc(&d, 148 * 8);
    for (uint32_t n : {256u, 128u}) {
        run("no swizzle, LBO 128, SBO 1024, K step 256 B (fpq_gemm.cu's layout)", 128, 1024, 0, 256, n, d);
        run("no swizzle, LBO 128, SBO 256, K step = rows/8 * 256 B (per-K-step tiles)", 128, 256, 0, n == 256 ? 8192 : 4096, n, d);
        run("no swizzle, LBO 1024 (k-chunk planes), SBO 128 (row groups contiguous)", n * 16, 128, 0, n * 32, n, d);
        run("128-byte swizzle, SBO 1024, K step 32 B", 16, 1024, 2, 32, n, d);
    }
    return 0;
}
