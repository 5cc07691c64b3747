// Instruction-throughput micro-benchmark for the packed fp16 quantizer's candidate instructions on sm_100a
// (development aid; build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fpqvar_b200/variants/ubench tools/ubench.cu).
// Every test runs CHAINS independent dependency chains of one instruction per thread, 32 warps per SM on every SM, and
// reports warp-instructions per clock per SM (4.0 = one per scheduler per clock).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

constexpr int CHAINS = 8;
constexpr int ITERS = 4096;

#define DEF_TEST(NAME, BODY)                                                                       \
    __global__ void __launch_bounds__(1024) NAME(uint32_t* out, uint32_t seed, long long* cyc) {   \
        uint32_t r[CHAINS];                                                                        \
        float f[CHAINS];                                                                           \
        uint64_t d[CHAINS];                                                                        \
        for (int c = 0; c < CHAINS; ++c) { r[c] = seed + threadIdx.x * 31 + c; f[c] = float(r[c] & 1023) * 0.001f + 0.5f; d[c] = (uint64_t(__float_as_uint(f[c])) << 32) | __float_as_uint(f[c]); } \
        const uint32_t k = seed | 0x3c003c00u;                                                     \
        const float kf = 1.0001f;                                                                  \
        const uint64_t kd = (uint64_t(__float_as_uint(kf)) << 32) | __float_as_uint(kf);            \
        (void)k; (void)kf; (void)kd;                                                               \
        const long long t0 = clock64();                                                            \
        for (int i = 0; i < ITERS; ++i) {                                                          \
            _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) { BODY }                            \
        }                                                                                          \
        const long long t1 = clock64();                                                            \
        uint32_t acc = 0;                                                                          \
        for (int c = 0; c < CHAINS; ++c) acc ^= r[c] ^ __float_as_uint(f[c]) ^ uint32_t(d[c]) ^ uint32_t(d[c] >> 32); \
        if (acc == 0x12345u) out[threadIdx.x] = acc;                                               \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                           \
    }

DEF_TEST(t_lop3, asm volatile("lop3.b32 %0, %0, %1, %0, 0x96;" : "+r"(r[c]) : "r"(k));)
DEF_TEST(t_prmt, asm volatile("prmt.b32 %0, %0, %1, 0xBB99;" : "+r"(r[c]) : "r"(k));)
DEF_TEST(t_imad, asm volatile("mad.lo.s32 %0, %0, %1, %0;" : "+r"(r[c]) : "r"(k));)
DEF_TEST(t_ffma, asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(f[c]) : "f"(kf));)
DEF_TEST(t_ffma2, asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(d[c]) : "l"(kd));)
DEF_TEST(t_fmul2, asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(d[c]) : "l"(kd));)
DEF_TEST(t_fadd2, asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[c]) : "l"(kd));)
DEF_TEST(t_fmnmx, asm volatile("max.f32 %0, %0, %1;" : "+f"(f[c]) : "f"(kf));)
DEF_TEST(t_hfma2, asm volatile("fma.rn.f16x2 %0, %0, %1, %0;" : "+r"(r[c]) : "r"(k));)
DEF_TEST(t_hadd2, asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(r[c]) : "r"(k));)
DEF_TEST(t_hmnmx2, asm volatile("max.NaN.f16x2 %0, %0, %1;" : "+r"(r[c]) : "r"(k));)
DEF_TEST(t_vmaxu2, r[c] = __vmaxu2(r[c], k);)
DEF_TEST(t_fhadd, { float o; asm volatile("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; add.rn.f32.f16 %0, lo, %2; }" : "=f"(o) : "r"(r[c]), "f"(kf)); r[c] = __float_as_uint(o); })
DEF_TEST(t_widen, { float o; asm volatile("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo; }" : "=f"(o) : "r"(r[c])); r[c] = __float_as_uint(o); })
DEF_TEST(t_f2fp_pack, asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r[c]) : "f"(__uint_as_float(r[c])), "f"(kf));)
DEF_TEST(t_f2fp_e2m1, { asm volatile("{ .reg .b8 t; .reg .b32 u; cvt.rn.satfinite.e2m1x2.f32 t, %1, %2; cvt.u32.u8 u, t; xor.b32 %0, %1, u; }" : "=r"(r[c]) : "f"(__uint_as_float(r[c])), "f"(kf)); })
DEF_TEST(t_e2m1_unpack, { asm volatile("{ .reg .b8 t; cvt.u8.u32 t, %0; cvt.rn.f16x2.e2m1x2 %0, t; }" : "+r"(r[c])); })
DEF_TEST(t_e2m1_roundtrip, { asm volatile("{ .reg .b8 t; cvt.rn.satfinite.e2m1x2.f32 t, %1, %2; cvt.rn.f16x2.e2m1x2 %0, t; }" : "=r"(r[c]) : "f"(__uint_as_float(r[c])), "f"(kf)); })
DEF_TEST(t_f2fp_e2m3, { asm volatile("{ .reg .b16 t; .reg .b32 u; cvt.rn.satfinite.e2m3x2.f32 t, %1, %2; cvt.u32.u16 u, t; xor.b32 %0, %1, u; }" : "=r"(r[c]) : "f"(__uint_as_float(r[c])), "f"(kf)); })
DEF_TEST(t_shfl, r[c] = __shfl_xor_sync(0xffffffffu, r[c], 1);)
DEF_TEST(t_mufu_rcp, asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[c]));)

typedef void (*kern_t)(uint32_t*, uint32_t, long long*);
struct Test { const char* name; kern_t k; int extra; };   // extra: helper instructions per body that are not the one measured

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, sizeof(long long) * sms * 2);
    Test tests[] = {
        {"LOP3", t_lop3, 0}, {"PRMT", t_prmt, 0}, {"IMAD", t_imad, 0}, {"FFMA", t_ffma, 0}, {"FFMA2", t_ffma2, 0}, {"FMUL2", t_fmul2, 0},
        {"FADD2", t_fadd2, 0}, {"FMNMX", t_fmnmx, 0}, {"HFMA2", t_hfma2, 0}, {"HADD2", t_hadd2, 0}, {"HMNMX2.NAN", t_hmnmx2, 0},
        {"VIMNMX.U16x2", t_vmaxu2, 0}, {"FHADD (f32 = f16 + f32)", t_fhadd, 0}, {"HADD2.F32 (widen)", t_widen, 0},
        {"F2FP.F16.F32.PACK_AB", t_f2fp_pack, 0}, {"F2FP.E2M1 pack (+1 LOP3)", t_f2fp_e2m1, 1}, {"F2FP.E2M1 unpack", t_e2m1_unpack, 0},
        {"F2FP.E2M1 pack+unpack (2 instr)", t_e2m1_roundtrip, 1}, {"F2FP.E2M3 pack (+1 LOP3)", t_f2fp_e2m3, 1}, {"SHFL.BFLY", t_shfl, 0},
        {"MUFU.RCP", t_mufu_rcp, 0},
    };
    printf("%-36s %14s %14s\n", "instruction", "warp-instr/clk/SM", "(body instr/clk/SM incl. helpers)");
    for (auto& t : tests) {
        t.k<<<sms, 1024>>>(out, 1u, cyc);
        cudaDeviceSynchronize();
        t.k<<<sms, 1024>>>(out, 1u, cyc);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s: CUDA error\n", t.name); return 1; }
        long long h[1024];
        cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        double mean = 0;
        for (int i = 0; i < sms; ++i) mean += double(h[i]);
        mean /= sms;
        const double bodies = double(ITERS) * CHAINS * 32.0;            // warp-level bodies per SM (32 warps)
        printf("%-36s %14.3f %14.3f\n", t.name, bodies / mean, bodies * (1 + t.extra) / mean);
    }
    return 0;
}
