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
    return 0;
}
