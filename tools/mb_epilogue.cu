// Microbenchmark (development aid): cycles per call of the f16x3 hidden-layer epilogue of kernels_tc3.cuh in isolation,
// with one tile (8 warps) and two tiles (16 warps) resident on one SM.  If two tiles take twice as long as one, the CUDA
// cores are saturated by this instruction mix; if they take the same, co-residency hides the latencies.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../include -o mb_epilogue mb_epilogue.cu
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "flowchain_b200.h"
#include "../flowchain_iccv2023_b200/csrc/kernels_fp32.cuh"
#include "../flowchain_iccv2023_b200/csrc/kernels_tc.cuh"
#include "../flowchain_iccv2023_b200/csrc/kernels_tc3.cuh"
using namespace fc; using namespace fc::tc; using namespace fc::t2; using namespace fc::t3;
__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }
// ---- experimental variant: NV (16 or 32) contiguous columns per load, every stage applied to all NV values before the
// next one (breadth-first source order) ----
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
template <int ACT, int NV, bool NOINL>
__device__ __forceinline__ void epi_bf_body(uint32_t acc, uint32_t ohi, uint32_t olo, int part, uint32_t koff) {
    constexpr int COLS = 32;                      // W = 64, two parts
    const int c0 = part * COLS;
#pragma unroll
    for (int q = 0; q < COLS; q += NV) {
        float v[NV];
        if (NV == 32) tmem_ld32(acc + (uint32_t)(c0 + q), v);
        else tmem_ld16(acc + (uint32_t)(c0 + q), v);
        tmem_wait_ld();
        if (ACT == ACT_T) {
            float e[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) v[i] = fminf(v[i], 43.0f);
#pragma unroll
            for (int i = 0; i < NV; ++i) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[i]) : "f"(v[i]));
#pragma unroll
            for (int i = 0; i < NV; ++i) e[i] += 1.0f;
            float r[NV / 2];
#pragma unroll
            for (int i = 0; i < NV / 2; ++i) r[i] = e[2 * i] * e[2 * i + 1];
#pragma unroll
            for (int i = 0; i < NV / 2; ++i) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r[i]) : "f"(r[i]));
#pragma unroll
            for (int i = 0; i < NV / 2; ++i) r[i] *= -2.0f;
#pragma unroll
            for (int i = 0; i < NV / 2; ++i) { v[2 * i] = fmaf(e[2 * i + 1], r[i], 1.0f); v[2 * i + 1] = fmaf(e[2 * i], r[i], 1.0f); }
        }
        float h[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) h[i] = __uint_as_float(__float_as_uint(v[i]) & 0xFFFFE000u);
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] -= h[i];
        uint32_t ph[NV / 2], pl[NV / 2];
#pragma unroll
        for (int i = 0; i < NV / 2; ++i) {
            if (ACT == ACT_R) {
                asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(ph[i]) : "f"(h[2 * i + 1]), "f"(h[2 * i]));
                asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(pl[i]) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
            } else {
                asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(ph[i]) : "f"(h[2 * i + 1]), "f"(h[2 * i]));
                asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pl[i]) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
            }
        }
#pragma unroll
        for (int i = 0; i < NV / 8; ++i) {
            const uint32_t off = koff + (uint32_t)((c0 + q + 8 * i) >> 3) * 128u;
            sts128(ohi + off, ph + 4 * i);
            sts128(olo + off, pl + 4 * i);
        }
    }
}
template <int ACT, int NV>
__device__ __noinline__ void epi_bf(uint32_t acc, uint32_t ohi, uint32_t olo, int part, uint32_t koff) { epi_bf_body<ACT, NV, true>(acc, ohi, olo, part, koff); }

template <int ACT, int NV>
__global__ void __launch_bounds__(512, 1) kbf(long long* out, int iters) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = warp >> 3, part = (warp >> 2) & 1, row = (warp & 3) * 32 + lane;
    if (warp == 0) tmem_alloc(&tslot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = tslot + (uint32_t)g * 256 + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t koff = (uint32_t)((row >> 3) * 8 * 128 + (row & 7) * 16);
    const uint32_t ohi = smem_u32(sm) + (uint32_t)g * 2u * 128u * 80u * 2u, olo = ohi + 128u * 80u * 2u;
    __syncthreads();
    const long long t0 = clk();
    for (int i = 0; i < iters; ++i) epi_bf<ACT, NV>(tb + kT3Acc0, ohi, olo, part, koff);
    const long long t1 = clk();
    __syncthreads();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / iters;
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tslot, 512);
}

template <int ACT, bool DOT, bool WRITE, bool BIAS, int NB>
__global__ void __launch_bounds__(512, 1) k(long long* out, int iters) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint32_t tslot;
    __shared__ __align__(16) float bias[512];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = warp >> 3, part = (warp >> 2) & 1, row = (warp & 3) * 32 + lane;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) bias[i] = 0.01f * i;
    if (warp == 0) tmem_alloc(&tslot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = tslot + (uint32_t)g * 256 + ((uint32_t)((warp & 3) * 32) << 16);
    const int W = 16 * NB;
    const uint32_t koff = (uint32_t)((row >> 3) * (W >> 3) * 128 + (row & 7) * 16);
    const uint32_t ohi = smem_u32(sm) + (uint32_t)g * 2u * 128u * 80u * 2u, olo = ohi + 128u * 80u * 2u;
    float d = 0;
    __syncthreads();
    const long long t0 = clk();
    for (int i = 0; i < iters; ++i)
        d += epi_hidden3_nb<3, ACT, DOT, WRITE, BIAS, NB>(tb + kT3Acc0, ohi, olo, smem_u32(bias), smem_u32(bias) + 1024, part, koff);
    const long long t1 = clk();
    __syncthreads();
    if (threadIdx.x == 0) { out[0] = (t1 - t0) / iters; out[1] = (long long)d; }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tslot, 512);
}
int main() {
    long long *d, h[8]; cudaMalloc(&d, 64);
    const int smem = 4 * 128 * 80 * 2;
#define RUN(ACT, DOT, WR, BIAS, NB, name)                                                                              \
    for (int warps = 8; warps <= 16; warps += 8) {                                                                      \
        cudaFuncSetAttribute(k<ACT, DOT, WR, BIAS, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);             \
        k<ACT, DOT, WR, BIAS, NB><<<1, warps * 32, smem>>>(d, 400);                                                     \
        cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);                          \
        printf("%-28s W=%d %2d warps: %5lld cycles/call (%s)\n", name, 16 * NB, warps, h[0], cudaGetErrorString(e));    \
    }
    RUN(ACT_T, false, true, false, 4, "tanh write, folded bias");
    RUN(ACT_R, false, true, false, 4, "relu write, folded bias");
    RUN(ACT_T, true, false, false, 4, "tanh dot, folded bias");
    RUN(ACT_R, false, true, true, 5, "relu write, bias add");
#define RUNBF(ACT, NV, name)                                                                                           \
    for (int warps = 8; warps <= 16; warps += 8) {                                                                      \
        cudaFuncSetAttribute(kbf<ACT, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                          \
        kbf<ACT, NV><<<1, warps * 32, smem>>>(d, 400);                                                                  \
        cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);                          \
        printf("%-28s W=64 %2d warps: %5lld cycles/call (%s)\n", name, warps, h[0], cudaGetErrorString(e));             \
    }
    RUNBF(ACT_T, 16, "tanh breadth-first x16");
    RUNBF(ACT_T, 32, "tanh breadth-first x32");
    RUNBF(ACT_R, 16, "relu breadth-first x16");
    RUNBF(ACT_R, 32, "relu breadth-first x32");
    return 0;
}
