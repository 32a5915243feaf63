// tc_host.cuh -- host-side helpers of the tensor-core paths (fp16/bf16 rounding, canonical K-major packing, blob
// pointers per step) and the tcgen05 self-test kernels (fc_tc_gemm_selftest).
#pragma once
#include <cuda_fp16.h>
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

// TS variant of the self-test: A lives in TENSOR MEMORY as packed fp16 pairs written with tcgen05.st by the
// thread that owns the row (exactly how the fused kernel hands activations to the next layer); B (fp16) in smem.
__global__ void __launch_bounds__(128)
k_tc_selftest_ts(const float* __restrict__ gAf, const uint8_t* __restrict__ gB, float* __restrict__ D, int N, int K, int Kp, int ncols) {
    using namespace tc;
    extern __shared__ __align__(128) uint8_t smem_raw2[];
    uint8_t* sB = smem_raw2;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + N * Kp * 2);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
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
    const uint32_t tlane = taddr + ((uint32_t)(warp * 32) << 16);
    const uint32_t a_col = 256;              // A operand at columns [256, 256 + Kp/2)
    if (tid == 0) {
        mbar_expect_tx(&bars[0], (uint32_t)(N * Kp * 2));
        bulk_load(sB, gB, (uint32_t)(N * Kp * 2), &bars[0]);
    }
    for (int k0 = 0; k0 < Kp; k0 += 16) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int ka = k0 + 2 * i, kb = ka + 1;
            const float lo = ka < K ? gAf[(size_t)tid * K + ka] : 0.0f, hi = kb < K ? gAf[(size_t)tid * K + kb] : 0.0f;
            __half2 h = __floats2half2_rn(lo, hi);
            v[i] = __uint_as_float(*reinterpret_cast<uint32_t*>(&h));
        }
        tmem_st8(tlane + a_col + (uint32_t)(k0 >> 1), v);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_f16(128, N);
        const uint32_t sbo = (uint32_t)(Kp >> 3) * 128u;
        for (int ks = 0; ks < (Kp >> 4); ++ks)
            umma_ts(taddr, taddr + a_col + (uint32_t)(ks * 8), umma_desc(smem_u32(sB) + ks * 256, 128, sbo), idesc, ks > 0 ? 1u : 0u);
        umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 8) {
        float v[8];
        tmem_ld8(tlane + (uint32_t)c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) D[(size_t)tid * N + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(taddr, (uint32_t)ncols);
}

inline uint16_t f2h(float f) {   // fp32 -> fp16 round-to-nearest-even (host)
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (x > 0x7f800000u ? 0x200u : 0));
    if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);          // overflow -> inf
    if (x < 0x33000001u) return (uint16_t)sign;                         // underflow -> 0
    int exp = (int)(x >> 23) - 127;
    uint32_t man = (x & 0x7fffffu) | 0x800000u;
    int shift = exp < -14 ? (13 + (-14 - exp)) : 13;
    uint32_t half_man = man >> shift;
    const uint32_t rem = man & ((1u << shift) - 1), halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (half_man & 1))) half_man++;
    uint32_t h;
    if (exp < -14) h = half_man;                                         // subnormal (may round up into normal)
    else h = ((uint32_t)(exp + 15) << 10) + (half_man - 0x400u);
    return (uint16_t)(sign | h);
}

inline float h2f(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ffu, x;
    if (exp == 0) {
        if (man == 0) x = sign;
        else {
            int e = -1;
            do { man <<= 1; ++e; } while (!(man & 0x400u));
            x = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) x = sign | 0x7f800000u | (man << 13);
    else x = sign | ((exp + 112) << 23) | (man << 13);
    float f;
    memcpy(&f, &x, 4);
    return f;
}

// [rows x K] fp32 row-major -> canonical K-major fp16 image; part 0 = hi = fp16(w), part 1 = lo = fp16(w - hi)
inline void pack_kmajor_f16(uint8_t* dst, const float* src, int rows, int K, int rows_pad, int Kp, int part) {
    memset(dst, 0, (size_t)rows_pad * Kp * 2);
    for (int r = 0; r < rows; ++r)
        for (int k = 0; k < K; ++k) {
            const float w = src[(size_t)r * K + k];
            uint16_t b = f2h(w);
            if (part == 1) b = f2h(w - h2f(b));
            memcpy(dst + tc::kmajor_offset(r, k, Kp), &b, 2);
        }
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

inline int tc_selftest(const float* A, const float* B, float* D, int N, int K, int device, std::string* msg, int ts_mode = 0) {
    if (N % 16 || N < 16 || N > 256 || K < 1 || K > 256) { *msg = "selftest: need N in [16,256] multiple of 16, K in [1,256]"; return 1; }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { *msg = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return 1; }
    const int Kp = (K + 15) / 16 * 16;
    std::vector<uint8_t> hA((size_t)128 * Kp * 2), hB((size_t)N * Kp * 2);
    pack_kmajor(hA.data(), A, 128, K, 128, Kp);
    if (ts_mode) pack_kmajor_f16(hB.data(), B, N, K, N, Kp, 0);
    else pack_kmajor(hB.data(), B, N, K, N, Kp);
    float* dAf = nullptr;
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
        if (ts_mode) {
            if ((e = cudaMalloc(&dAf, (size_t)128 * K * 4)) != cudaSuccess) break;
            if ((e = cudaMemcpy(dAf, A, (size_t)128 * K * 4, cudaMemcpyHostToDevice)) != cudaSuccess) break;
            if ((e = cudaFuncSetAttribute(k_tc_selftest_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) break;
            k_tc_selftest_ts<<<1, 128, smem>>>(dAf, dB, dD, N, K, Kp, 512);
        } else
        k_tc_selftest<<<1, 128, smem>>>(dA, dB, dD, N, Kp, ncols);
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        if ((e = cudaDeviceSynchronize()) != cudaSuccess) break;
        if ((e = cudaMemcpy(D, dD, (size_t)128 * N * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        rc = 0;
    } while (0);
    if (rc) *msg = std::string("selftest CUDA error: ") + cudaGetErrorString(e);
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dAf);
    return rc;
}

// host pointers into the canonical blob, per step
struct TcStepSrc {
    const float* cpl_W[kMaxBlocks][2][kMaxHidden + 2];
    const float* cpl_b[kMaxBlocks][2][kMaxHidden + 2];
    const float* dens_W[2][2][3];
    const float* dens_b[2][2][3];
};

inline int pad16i(int v) { return (v + 15) / 16 * 16; }

}  // namespace fc
