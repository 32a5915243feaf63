// mb_wake.cu -- microbenchmark (development aid): latency from mbarrier.arrive in one warp to the waiting warp running
// again, for the wait flavours: try_wait with a suspend-time hint (what mbar_wait uses), plain try_wait, test_wait spin.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mb_wake_bin tools/mb_wake.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../flowchain_iccv2023_b200/csrc/tc_common.cuh"
using namespace fc::tc;

__device__ __forceinline__ bool try_wait_plain(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool try_wait_hint(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
template <int MODE>
__device__ __forceinline__ void wait_mode(uint64_t* bar, uint32_t parity) {
    if (MODE == 0) { while (!try_wait_hint(bar, parity)) {} }
    else if (MODE == 1) { while (!try_wait_plain(bar, parity)) {} }
    else { while (!test_wait(bar, parity)) {} }
}
// warp 0 waits on `go` (flavour MODE), warp 1 arrives after a delay; `back` (spin) returns the turn.  `busy` extra warps run FMAs.
template <int MODE>
__global__ void k(int reps, int delay, long long* out) {
    __shared__ uint64_t go, back;
    __shared__ long long t_arr;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&go, 1); mbar_init(&back, 1); mbar_fence_init(); stop = 0; }
    __syncthreads();
    if (warp == 0) {
        long long sum = 0, mx = 0;
        for (int r = 0; r < reps; ++r) {
            wait_mode<MODE>(&go, r & 1);
            const long long t = clock64();
            const long long d = t - *(volatile long long*)&t_arr;
            sum += d; mx = d > mx ? d : mx;
            __syncwarp();
            if (lane == 0) mbar_arrive(&back);
        }
        if (lane == 0) { out[0] = sum; out[1] = mx; stop = 1; }
    } else if (warp == 1) {
        for (int r = 0; r < reps; ++r) {
            if (r > 0) { while (!test_wait(&back, (r - 1) & 1)) {} }
            const long long t0 = clock64();
            while (clock64() - t0 < delay) {}
            if (lane == 0) { *(volatile long long*)&t_arr = clock64(); __threadfence_block(); mbar_arrive(&go); }
            __syncwarp();
        }
    } else {
        float f = threadIdx.x;
        while (!stop) {
#pragma unroll
            for (int i = 0; i < 64; ++i) f = fmaf(f, 1.0001f, 0.5f);
        }
        if (f == 123.f) out[2] = 1;
    }
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    const int reps = 2000;
    printf("# arrive -> waiter running again, cycles (avg, max); delay = how long the waiter has been waiting\n");
    for (int threads : {64, 640}) for (int delay : {100, 2000}) {
        long long h[3][2];
        k<0><<<1, threads>>>(reps, delay, d); cudaDeviceSynchronize(); cudaMemcpy(h[0], d, 16, cudaMemcpyDeviceToHost);
        k<1><<<1, threads>>>(reps, delay, d); cudaDeviceSynchronize(); cudaMemcpy(h[1], d, 16, cudaMemcpyDeviceToHost);
        k<2><<<1, threads>>>(reps, delay, d); cudaDeviceSynchronize(); cudaMemcpy(h[2], d, 16, cudaMemcpyDeviceToHost);
        printf("threads %3d delay %4d : try_wait+hint %6.1f (max %5lld) | try_wait %6.1f (max %5lld) | test_wait spin %6.1f (max %5lld)   %s\n", threads, delay,
               (double)h[0][0] / reps, h[0][1], (double)h[1][0] / reps, h[1][1], (double)h[2][0] / reps, h[2][1], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
