// tcgen05.ld (tensor memory -> registers) throughput on sm_100a by shape and by number of reading warps (development aid for the
// low-bit GEMM's epilogue; build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fpqvar_b200/variants/tmem_ldbench tools/tmem_ldbench.cu).
// One CTA per SM, 512 columns allocated; every warp loops over loads of its own lane quarter; prints bytes per clock per SM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

#define R4(b) "%" #b ", %" #b "+1"
template <int SHAPE>
__device__ __forceinline__ uint32_t do_ld(uint32_t taddr) {
    uint32_t v[64];
    if constexpr (SHAPE == 0) {        // 32x32b.x16: 32 lanes x 16 columns = 2 KB
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        return v[0] ^ v[15];
    } else if constexpr (SHAPE == 1) { // 32x32b.x64: 32 lanes x 64 columns = 8 KB
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                     "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                       "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                       "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
                       "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]),
                       "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]),
                       "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        return v[0] ^ v[63];
    } else if constexpr (SHAPE == 2) { // 16x256b.x4: 16 lanes x 1024 bits = 2 KB (16 registers per thread)
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        return v[0] ^ v[15];
    } else if constexpr (SHAPE == 3) { // 16x128b.x8: 16 lanes x 1024 bits = 2 KB (16 registers per thread)
        asm volatile("tcgen05.ld.sync.aligned.16x128b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        return v[0] ^ v[15];
    } else {                            // 32x32b.x16 twice, one wait (two loads in flight)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
                       "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr + 16));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        return v[0] ^ v[31];
    }
}

template <int SHAPE>
__global__ void __launch_bounds__(512) bench(int iters, unsigned long long* cycles, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) acc ^= do_ld<SHAPE>(base + ((i * 64) & 255));
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

template <int SHAPE>
void run(const char* name, int bytes_per_ld, unsigned long long* d_cyc, uint32_t* d_sink) {
    const int iters = 4096;
    for (int warps : {1, 4, 8, 16}) {
        bench<SHAPE><<<148, warps * 32>>>(iters, d_cyc, d_sink);
        bench<SHAPE><<<148, warps * 32>>>(iters, d_cyc, d_sink);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long h[148];
        cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
        unsigned long long mx = 0;
        for (auto c : h) mx = c > mx ? c : mx;
        printf("%-28s warps %2d: %7.1f clk per load per warp, %7.1f B/clk per SM  (%s)\n", name, warps, double(mx) / iters,
               double(bytes_per_ld) * warps * iters / double(mx), cudaGetErrorString(e));
    }
}

int main() {
    unsigned long long* d_cyc;
    uint32_t* d_sink;
    cudaMalloc(&d_cyc, 148 * sizeof(unsigned long long));
    cudaMalloc(&d_sink, 64);
    run<0>("32x32b.x16 (2 KB)", 2048, d_cyc, d_sink);
    run<4>("32x32b.x16 x2 in flight (4 KB)", 4096, d_cyc, d_sink);
    run<1>("32x32b.x64 (8 KB)", 8192, d_cyc, d_sink);
    run<2>("16x256b.x4 (2 KB)", 2048, d_cyc, d_sink);
    run<3>("16x128b.x8 (2 KB)", 2048, d_cyc, d_sink);
    return 0;
}
