// Which (lane, column) of tensor memory lands in which (thread, register) of tcgen05.ld.16x256b.x2 / 16x128b.x2 (development aid;
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fpqvar_b200/variants/tmem_layout_probe tools/tmem_layout_probe.cu).
// Tensor memory is filled through tcgen05.st.32x32b (thread = lane, register = column) with lane * 1000 + column.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void __launch_bounds__(128) probe(uint32_t* out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + (uint32_t(warp * 32) << 16);
    uint32_t v[16];
    for (int c = 0; c < 16; ++c) v[c] = uint32_t((warp * 32 + lane) * 1000 + c);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(base),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                 "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[8];
    // 16x256b.x2: 16 lanes x 16 columns, 8 registers per thread; second half of the warp's lanes at lane offset 16
    for (int h = 0; h < 2; ++h) {
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(base + (uint32_t(h * 16) << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) out[((0 * 2 + h) * 128 + threadIdx.x) * 8 + i] = r[i];
    }
    // 16x128b.x2: 16 lanes x 8 columns, 4 registers per thread
    for (int h = 0; h < 2; ++h) {
        asm volatile("tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(base + (uint32_t(h * 16) << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 4; ++i) out[((1 * 2 + h) * 128 + threadIdx.x) * 8 + i] = r[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(32) : "memory");
}

int main() {
    uint32_t* d;
    cudaMalloc(&d, 4 * 128 * 8 * 4);
    cudaMemset(d, 0xff, 4 * 128 * 8 * 4);
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    static uint32_t h[4 * 128 * 8];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[2] = {"16x256b.x2", "16x128b.x2"};
    for (int s = 0; s < 2; ++s)
        for (int hh = 0; hh < 2; ++hh) {
            printf("%s, lane offset %d: thread -> (lane,col) per register  [warp 1 shown: lanes 32..63]\n", names[s], hh * 16);
            for (int t = 32; t < 64; ++t) {
                printf("  t%2d:", t - 32);
                for (int i = 0; i < (s == 0 ? 8 : 4); ++i) {
                    const uint32_t v = h[((s * 2 + hh) * 128 + t) * 8 + i];
                    printf(" (%u,%u)", v / 1000, v % 1000);
                }
                printf("\n");
            }
        }
    return 0;
}
