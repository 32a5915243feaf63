// kernels_tc.cuh -- bf16 tensor-core (tcgen05 / TMEM) path of the fused CIF step.  (stub: filled in next)
#pragma once
#include <cuda_runtime.h>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/flowchain_b200.h"
#include "plan.h"
#include "tc_common.cuh"

namespace fc {

// ---------------------------------------------------------------------------------------------
// tcgen05 self-test: D[128 x N] = A[128 x K] * B[N x K]^T with bf16 operands / fp32 accumulate,
// exercising exactly the helpers the fused kernel uses (generic-proxy A tile + fence, bulk-copy B
// tile, TMEM alloc, MMA issue, commit -> mbarrier, tcgen05.ld epilogue).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_tc_selftest(const uint8_t* __restrict__ gA, const uint8_t* __restrict__ gB, float* __restrict__ D, int N, int Kp, int ncols) {
    using namespace tc;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* sA = smem_raw;
    uint8_t* sB = sA + 128 * Kp * 2;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + N * Kp * 2);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tslot, (uint32_t)ncols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = *tslot;
    for (int i = tid; i < 128 * Kp * 2 / 16; i += 128)
        reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(gA)[i];
    fence_async_smem();
    if (tid == 0) {
        mbar_expect_tx(&bars[0], (uint32_t)(N * Kp * 2));
        bulk_load(sB, gB, (uint32_t)(N * Kp * 2), &bars[0]);
    }
    __syncthreads();
    if (tid == 0) {
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        umma_tile(taddr, smem_u32(sA), smem_u32(sB), Kp, N, false);
        umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 8) {
        float v[8];
        tmem_ld8(taddr + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) D[(size_t)(warp * 32 + lane) * N + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(taddr, (uint32_t)ncols);
}

inline uint16_t f2bf(float f) {   // round-to-nearest-even, host side
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

// [rows x K] fp32 row-major -> canonical K-major bf16 image ([rows_pad x Kp], zero padded)
inline void pack_kmajor(uint8_t* dst, const float* src, int rows, int K, int rows_pad, int Kp) {
    memset(dst, 0, (size_t)rows_pad * Kp * 2);
    for (int r = 0; r < rows; ++r)
        for (int k = 0; k < K; ++k) {
            const uint16_t b = f2bf(src[(size_t)r * K + k]);
            memcpy(dst + tc::kmajor_offset(r, k, Kp), &b, 2);
        }
}

inline int tc_selftest(const float* A, const float* B, float* D, int N, int K, int device, std::string* msg) {
    if (N % 16 || N < 16 || N > 256 || K < 1 || K > 256) { *msg = "selftest: need N in [16,256] multiple of 16, K in [1,256]"; return 1; }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { *msg = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return 1; }
    const int Kp = (K + 15) / 16 * 16;
    std::vector<uint8_t> hA((size_t)128 * Kp * 2), hB((size_t)N * Kp * 2);
    pack_kmajor(hA.data(), A, 128, K, 128, Kp);
    pack_kmajor(hB.data(), B, N, K, N, Kp);
    uint8_t *dA = nullptr, *dB = nullptr;
    float* dD = nullptr;
    int rc = 1;
    int ncols = 32;
    while (ncols < N) ncols *= 2;
    const size_t smem = hA.size() + hB.size() + 64;
    do {
        if ((e = cudaMalloc(&dA, hA.size())) != cudaSuccess) break;
        if ((e = cudaMalloc(&dB, hB.size())) != cudaSuccess) break;
        if ((e = cudaMalloc(&dD, (size_t)128 * N * 4)) != cudaSuccess) break;
        if ((e = cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaFuncSetAttribute(k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) break;
        k_tc_selftest<<<1, 128, smem>>>(dA, dB, dD, N, Kp, ncols);
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        if ((e = cudaDeviceSynchronize()) != cudaSuccess) break;
        if ((e = cudaMemcpy(D, dD, (size_t)128 * N * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        rc = 0;
    } while (0);
    if (rc) *msg = std::string("selftest CUDA error: ") + cudaGetErrorString(e);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return rc;
}

// host pointers into the canonical blob, per step
struct TcStepSrc {
    const float* cpl_W[kMaxBlocks][2][kMaxHidden + 2];
    const float* cpl_b[kMaxBlocks][2][kMaxHidden + 2];
    const float* dens_W[2][2][3];
    const float* dens_b[2][2][3];
};

inline void tc_pack(const fc_model_desc&, std::vector<StepPlan>&, const std::vector<TcStepSrc>&, std::vector<unsigned char>& img) {
    img.clear();
}

inline int tc_launch_separate(const fc_model_desc&, const StepPlan*, const unsigned char*, size_t, const std::vector<StepPlan>&,
                              const RowArgs&, int, cudaStream_t, uint64_t*, std::string* msg) {
    *msg = "the bf16 tensor-core path is not built in this library";
    return 1;
}

}  // namespace fc
