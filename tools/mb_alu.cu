// Microbenchmark (development aid): issue rate of the instructions the f16x3 operand split is made of, per SM per clock,
// with 16 warps resident (4 per scheduler) and 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mb_alu_bin tools/mb_alu.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int kIters = 4096, kChains = 8;
template <int OP>
__global__ void __launch_bounds__(512) k(float* out, float seed, long long* cyc) {
    float a[kChains]; uint32_t p[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) { a[i] = seed + threadIdx.x * 1e-3f + i; p[i] = 0x3c003c00u + i; }
    long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            if (OP == 0) a[i] = fmaf(a[i], 1.0001f, seed);                                     // FFMA
            if (OP == 1) asm volatile("{.reg .b16 x, y; mov.b32 {x, y}, %1; fma.rn.f32.f16 %0, x, y, %0;}" : "+f"(a[i]) : "r"(p[i]));   // FHFMA
            if (OP == 2) asm volatile("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(p[i]) : "f"(a[i]), "f"(__uint_as_float(p[i])));   // F2FP pack
            if (OP == 3) a[i] = __uint_as_float(__float_as_uint(a[i]) & (0xFFFFE000u + it));     // LOP3
            if (OP == 4) a[i] = fmaxf(a[i], seed + it);                                          // FMNMX
            if (OP == 5) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));                // MUFU
            if (OP == 6) {                                                                       // old split: LOP3 + FADD + 2 x 0.5 F2FP per value
                const float h = __uint_as_float(__float_as_uint(a[i]) & 0xFFFFE000u);
                const float l = a[i] - h; uint32_t ph, pl;
                asm volatile("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(ph) : "f"(h), "f"(h));
                asm volatile("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(pl) : "f"(l), "f"(l));
                a[i] += __uint_as_float(ph ^ pl);
            }
            if (OP == 7) {                                                                       // new split: F2FP + 2 FHFMA + F2FP per 2 values
                uint32_t ph, pl; float l0, l1;
                asm volatile("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(ph) : "f"(a[i]), "f"(a[i]));
                asm volatile("{.reg .b16 x, y, m; mov.b32 {x, y}, %2; mov.b16 m, 0xBC00; fma.rn.f32.f16 %0, x, m, %3; fma.rn.f32.f16 %1, y, m, %3;}" : "=f"(l0), "=f"(l1) : "r"(ph), "f"(a[i]));
                asm volatile("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pl) : "f"(l1), "f"(l0));
                a[i] += __uint_as_float(ph ^ pl);
            }
        }
    }
    long long t1 = clock64();
    float s = 0; uint32_t q = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) { s += a[i]; q ^= p[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(q);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name, float* out, long long* cyc, double per) {
    k<OP><<<148, 512>>>(out, 0.5f, cyc); cudaDeviceSynchronize();
    k<OP><<<148, 512>>>(out, 0.5f, cyc); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double thread_ops = (double)kIters * kChains * per;
    printf("%-28s %8lld cycles  %6.1f thread-ops/clk/SM  (%.2f warp-instr/clk/SM)\n", name, c, thread_ops * 512 / c, thread_ops * 16 / c);
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
    run<0>("FFMA", out, cyc, 1); run<1>("FHFMA (fma.f32.f16)", out, cyc, 1); run<2>("F2FP pack", out, cyc, 1);
    run<3>("LOP3", out, cyc, 1); run<4>("FMNMX", out, cyc, 1); run<5>("MUFU.EX2", out, cyc, 1);
    run<6>("old split (per value)", out, cyc, 1); run<7>("new split (per 2 values)", out, cyc, 2);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
