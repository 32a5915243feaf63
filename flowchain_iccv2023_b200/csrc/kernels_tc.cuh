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

constexpr int kTcMaxChunks = 6 + kMaxBlocks * (kMaxHidden + 1);
constexpr int kTcThreads = 288;      // 2 compute groups x 4 warps + 1 weight-producer warp
constexpr int kTcSlots = 3;          // weight-chunk ring depth
constexpr int kTcAccCols = 256;      // TMEM columns per group: accumulators [0,192), fp32 copy of u [192,256)
constexpr int kTcUCol = 192;

// Per-step description of the bf16 image: every layer is one "chunk" = canonical K-major B tile(s) followed
// by an fp32 tail (biases, and for the last hidden coupling layer the surviving row of the H->2 Linear).
struct TcStep {
    int32_t C, c, Kin, Ku, HPd, CPd, Hc;
    int32_t n_blocks, n_hidden, nchunks, eps_off, pad0;
    int32_t chunk_off[kTcMaxChunks];     // byte offset into the image
    int32_t chunk_bytes[kTcMaxChunks];   // multiple of 16
    int32_t masked_dim[kMaxBlocks];
    float bn_scale[kMaxBlocks][2], bn_shift[kMaxBlocks][2], bn_ldj[kMaxBlocks];
    float std[2];
};

struct TcDims {
    int32_t Kin_max, Ku_max, W_max;      // operand tile widths (elements)
    int32_t slot_bytes;                  // ring slot size (max chunk)
    int32_t T, C0;
    int32_t supported;
};

inline int pad16i(int v) { return (v + 15) / 16 * 16; }

// Builds the bf16 image.  Leaves `img` empty (path unavailable) for shapes the tensor-core kernel does not
// support: H must be a multiple of 16 with 2H <= 256 and every accumulator must fit 192 TMEM columns.
inline void tc_pack(const fc_model_desc& d, const std::vector<StepPlan>& plans, const std::vector<TcStepSrc>& src,
                    std::vector<unsigned char>& img, std::vector<TcStep>& steps, TcDims& dm) {
    img.clear();
    steps.clear();
    memset(&dm, 0, sizeof dm);
    const int T = d.pred_len, H = d.hidden_size;
    const int cmax = d.cond_size + (T - 1) * 2 + 2;
    if (H % 16 != 0 || 2 * H > kTcUCol || 2 * pad16i(2 * cmax) > kTcUCol || pad16i(cmax) > 64) return;
    dm.T = T; dm.C0 = d.cond_size;
    steps.resize(T);
    std::vector<float> tmp;
    auto add_tile = [&](std::vector<unsigned char>& out, const float* W, int rows, int K, int rows_pad, int Kp, int row_off, std::vector<float>* acc) {
        (void)row_off; (void)acc;
        const size_t o = out.size();
        out.resize(o + (size_t)rows_pad * Kp * 2);
        pack_kmajor(out.data() + o, W, rows, K, rows_pad, Kp);
    };
    auto add_floats = [&](std::vector<unsigned char>& out, const float* p, int n, int n_pad) {
        const size_t o = out.size();
        out.resize(o + (size_t)n_pad * 4, 0);
        memcpy(out.data() + o, p, (size_t)n * 4);
    };
    for (int i = 0; i < T; ++i) {
        TcStep& ts = steps[i];
        memset(&ts, 0, sizeof ts);
        const StepPlan& sp = plans[i];
        const TcStepSrc& s = src[i];
        const int C = sp.C, c = sp.c;
        ts.C = C; ts.c = c; ts.Kin = pad16i(c); ts.Ku = pad16i(C + 2); ts.HPd = pad16i(2 * c); ts.CPd = pad16i(C); ts.Hc = H;
        ts.n_blocks = sp.n_blocks; ts.n_hidden = sp.n_hidden; ts.eps_off = sp.eps_off;
        for (int b = 0; b < sp.n_blocks; ++b) {
            ts.masked_dim[b] = sp.masked_dim[b];
            ts.bn_scale[b][0] = sp.bn_scale[b][0]; ts.bn_scale[b][1] = sp.bn_scale[b][1];
            ts.bn_shift[b][0] = sp.bn_shift[b][0]; ts.bn_shift[b][1] = sp.bn_shift[b][1];
            ts.bn_ldj[b] = sp.bn_ldj[b];
        }
        ts.std[0] = sp.std[0]; ts.std[1] = sp.std[1];
        dm.Kin_max = ts.Kin > dm.Kin_max ? ts.Kin : dm.Kin_max;
        dm.Ku_max = ts.Ku > dm.Ku_max ? ts.Ku : dm.Ku_max;
        const int w = ts.HPd > H ? ts.HPd : H;
        dm.W_max = w > dm.W_max ? w : dm.W_max;
        int nch = 0;
        auto begin_chunk = [&]() { while (img.size() % 128) img.push_back(0); ts.chunk_off[nch] = (int32_t)img.size(); };
        auto end_chunk = [&]() {
            while (img.size() % 16) img.push_back(0);
            ts.chunk_bytes[nch] = (int32_t)(img.size() - ts.chunk_off[nch]);
            dm.slot_bytes = ts.chunk_bytes[nch] > dm.slot_bytes ? ts.chunk_bytes[nch] : dm.slot_bytes;
            ++nch;
        };
        auto dens_chunks = [&](int which) {
            const int HPd = ts.HPd, CPd = ts.CPd, Kin = ts.Kin, h2 = 2 * c;
            // L1: rows [0,HPd) shift net, [HPd,2HPd) log-scale net; K = c
            begin_chunk();
            tmp.assign((size_t)2 * HPd * c, 0.0f);
            std::vector<float> bias(2 * HPd, 0.0f);
            for (int net = 0; net < 2; ++net) {
                memcpy(tmp.data() + (size_t)net * HPd * c, s.dens_W[which][net][0], (size_t)h2 * c * 4);
                memcpy(bias.data() + net * HPd, s.dens_b[which][net][0], (size_t)h2 * 4);
            }
            add_tile(img, tmp.data(), 2 * HPd, c, 2 * HPd, Kin, 0, nullptr);
            add_floats(img, bias.data(), 2 * HPd, 2 * HPd);
            end_chunk();
            // L2: two [HPd x HPd] tiles
            begin_chunk();
            bias.assign(2 * HPd, 0.0f);
            for (int net = 0; net < 2; ++net) {
                add_tile(img, s.dens_W[which][net][1], h2, h2, HPd, HPd, 0, nullptr);
                memcpy(bias.data() + net * HPd, s.dens_b[which][net][1], (size_t)h2 * 4);
            }
            add_floats(img, bias.data(), 2 * HPd, 2 * HPd);
            end_chunk();
            // L3: two [CPd x HPd] tiles
            begin_chunk();
            bias.assign(2 * CPd, 0.0f);
            for (int net = 0; net < 2; ++net) {
                add_tile(img, s.dens_W[which][net][2], C, h2, CPd, HPd, 0, nullptr);
                memcpy(bias.data() + net * CPd, s.dens_b[which][net][2], (size_t)C * 4);
            }
            add_floats(img, bias.data(), 2 * CPd, 2 * CPd);
            end_chunk();
        };
        dens_chunks(0);
        for (int b = 0; b < sp.n_blocks; ++b) {
            const int j = 1 - sp.masked_dim[b];
            // L1: rows [0,H) s_net, [H,2H) t_net; K = C + 2
            begin_chunk();
            tmp.assign((size_t)2 * H * c, 0.0f);
            std::vector<float> bias(2 * H, 0.0f);
            for (int net = 0; net < 2; ++net) {
                memcpy(tmp.data() + (size_t)net * H * c, s.cpl_W[b][net][0], (size_t)H * c * 4);
                memcpy(bias.data() + net * H, s.cpl_b[b][net][0], (size_t)H * 4);
            }
            add_tile(img, tmp.data(), 2 * H, c, 2 * H, ts.Ku, 0, nullptr);
            add_floats(img, bias.data(), 2 * H, 2 * H);
            end_chunk();
            for (int l = 1; l <= sp.n_hidden; ++l) {
                begin_chunk();
                for (int net = 0; net < 2; ++net) {
                    add_tile(img, s.cpl_W[b][net][l], H, H, H, H, 0, nullptr);
                    memcpy(bias.data() + net * H, s.cpl_b[b][net][l], (size_t)H * 4);
                }
                add_floats(img, bias.data(), 2 * H, 2 * H);
                end_chunk();
            }
            // the surviving row (1 - mask) of the final H -> 2 Linear of s_net / t_net: appended to the LAST chunk
            // of the block (L1 when n_hidden == 0)
            {
                std::vector<float> last(2 * H + 4, 0.0f);
                for (int net = 0; net < 2; ++net) {
                    memcpy(last.data() + net * H, s.cpl_W[b][net][sp.n_hidden + 1] + (size_t)j * H, (size_t)H * 4);
                    last[2 * H + net] = s.cpl_b[b][net][sp.n_hidden + 1][j];
                }
                --nch;                                  // re-open the last chunk
                add_floats(img, last.data(), 2 * H + 4, 2 * H + 4);
                end_chunk();
            }
        }
        dens_chunks(1);
        ts.nchunks = nch;
    }
    while (img.size() % 128) img.push_back(0);
    dm.slot_bytes = (dm.slot_bytes + 127) / 128 * 128;
    dm.supported = 1;
}

inline size_t tc_smem_bytes(const TcDims& dm) {
    const size_t per_group = (size_t)128 * 2 * (dm.Kin_max + dm.Ku_max + 2 * dm.W_max);
    return 1024 + 2 * per_group + (size_t)kTcSlots * dm.slot_bytes;
}

namespace tcdev {
using namespace tc;

struct Ctx {
    int g, tid, quad;                 // group, thread in group (= row of the tile), warp % 4
    uint8_t *A_in, *A_u, *A_h0, *A_h1;
    uint32_t tacc;                    // TMEM address of this group's accumulator block (lane field 0)
    uint32_t tld;                     // same with this warp's lane quadrant, for tcgen05.ld/st
    uint64_t *full, *empty, *mma_bar;
    uint8_t* slots;
    int slot_bytes;
    uint32_t ring_it, mma_cnt;
};

__device__ __forceinline__ void group_bar(const Ctx& x) { asm volatile("bar.sync %0, 128;" ::"r"(1 + x.g) : "memory"); }

__device__ __forceinline__ const uint8_t* chunk_wait(const Ctx& x) {
    const uint32_t slot = x.ring_it % kTcSlots;
    mbar_wait(&x.full[slot], (x.ring_it / kTcSlots) & 1u);
    return x.slots + (size_t)slot * x.slot_bytes;
}
// end of a layer: operand tiles written by this group and TMEM reads are complete -> sync the group,
// hand the weight slot back to the producer.
__device__ __forceinline__ void layer_end(Ctx& x) {
    tmem_wait_ld();
    fence_async_smem();
    tc_fence_before();
    group_bar(x);
    if (x.tid == 0) mbar_arrive(&x.empty[x.ring_it % kTcSlots]);
    x.ring_it++;
}
__device__ __forceinline__ void mma_wait(Ctx& x) {
    mbar_wait(x.mma_bar, x.mma_cnt & 1u);
    x.mma_cnt++;
    tc_fence_after();
}

enum { A_RELU = 0, A_TANH = 1 };

// Hidden-layer epilogue: acc[:, net*W + k] + bias -> activation -> bf16 operand tile of the next layer.
// With DOT the fp32 activations are also contracted with the surviving row of the final Linear.
template <int ACT0, int ACT1, bool DOT>
__device__ __forceinline__ void epi_hidden(const Ctx& x, int W, const float* bias, const float* wlast, float& dot0, float& dot1) {
#pragma unroll 1
    for (int net = 0; net < 2; ++net) {
        uint8_t* A = net ? x.A_h1 : x.A_h0;
        const int act = net ? ACT1 : ACT0;
        float dot = 0.0f;
#pragma unroll 1
        for (int c0 = 0; c0 < W; c0 += 16) {
            float v[16];
            tmem_ld16(x.tld + (uint32_t)(net * W + c0), v);
            tmem_wait_ld();
            const float4* bp = reinterpret_cast<const float4*>(bias + net * W + c0);
            uint32_t pk[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 b4 = bp[q];
                float a0 = v[4 * q] + b4.x, a1 = v[4 * q + 1] + b4.y, a2 = v[4 * q + 2] + b4.z, a3 = v[4 * q + 3] + b4.w;
                if (act == A_RELU) {
                    a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f);
                    pk[2 * q] = pack_bf16x2(a0, a1);
                    pk[2 * q + 1] = pack_bf16x2(a2, a3);
                } else if (DOT) {
                    asm("tanh.approx.f32 %0, %0;" : "+f"(a0));
                    asm("tanh.approx.f32 %0, %0;" : "+f"(a1));
                    asm("tanh.approx.f32 %0, %0;" : "+f"(a2));
                    asm("tanh.approx.f32 %0, %0;" : "+f"(a3));
                    pk[2 * q] = pack_bf16x2(a0, a1);
                    pk[2 * q + 1] = pack_bf16x2(a2, a3);
                } else {
                    pk[2 * q] = tanh_bf16x2(pack_bf16x2(a0, a1));
                    pk[2 * q + 1] = tanh_bf16x2(pack_bf16x2(a2, a3));
                }
                if (DOT) {
                    const float4 w4 = reinterpret_cast<const float4*>(wlast + net * W + c0)[q];
                    dot = fmaf(a0, w4.x, dot); dot = fmaf(a1, w4.y, dot); dot = fmaf(a2, w4.z, dot); dot = fmaf(a3, w4.w, dot);
                }
            }
            uint8_t* dst = A + kmajor_offset(x.tid, c0, W);
            *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(dst + 128) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        if (net == 0) dot0 = dot; else dot1 = dot;
    }
}

}  // namespace tcdev

// ---------------------------------------------------------------------------------------------
// a8 on the tensor cores: per-step density, one work item = (step, pair of 128-row tiles).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1)
k_separate_tc(const TcStep* __restrict__ steps, const uint8_t* __restrict__ img, RowArgs a, TcDims dm,
              long long pairs_per_step, long long total_items) {
    using namespace tc;
    using namespace tcdev;
    extern __shared__ __align__(128) uint8_t smem_tc[];
    uint8_t* smem = smem_tc;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + kTcSlots;
    uint64_t* mma_bar = empty + kTcSlots;          // [2]
    uint32_t* tslot = reinterpret_cast<uint32_t*>(mma_bar + 2);
    TcStep* cur = reinterpret_cast<TcStep*>(smem + 128);
    static_assert(sizeof(TcStep) <= 896, "TcStep must fit the header");
    const size_t tile_in = (size_t)128 * 2 * dm.Kin_max, tile_u = (size_t)128 * 2 * dm.Ku_max, tile_h = (size_t)128 * 2 * dm.W_max;
    const size_t per_group = tile_in + tile_u + 2 * tile_h;
    uint8_t* gbase = smem + 1024;
    uint8_t* slots = gbase + 2 * per_group;

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < kTcSlots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2); }
        mbar_init(&mma_bar[0], 1);
        mbar_init(&mma_bar[1], 1);
        mbar_fence_init();
    }
    if (warp == 8) tmem_alloc(tslot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tslot;

    Ctx x;
    x.g = warp >> 2; x.tid = tid & 127; x.quad = warp & 3;
    if (warp < 8) {
        uint8_t* gb = gbase + (size_t)x.g * per_group;
        x.A_in = gb; x.A_u = gb + tile_in; x.A_h0 = x.A_u + tile_u; x.A_h1 = x.A_h0 + tile_h;
        x.tacc = tmem_base + (uint32_t)(x.g * kTcAccCols);
        x.tld = x.tacc + ((uint32_t)(x.quad * 32) << 16);
        x.mma_bar = &mma_bar[x.g];
    }
    x.full = full; x.empty = empty; x.slots = slots; x.slot_bytes = dm.slot_bytes;
    x.ring_it = 0; x.mma_cnt = 0;
    uint32_t prod_it = 0;

    for (long long item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int step = dm.T - 1 - (int)(item / pairs_per_step);      // heavy (late) steps first
        const long long pair = item % pairs_per_step;
        __syncthreads();
        {
            const int n = sizeof(TcStep) / 4;
            const int* s = reinterpret_cast<const int*>(steps + step);
            int* d = reinterpret_cast<int*>(cur);
            for (int i = tid; i < n; i += kTcThreads) d[i] = __ldg(s + i);
        }
        __syncthreads();
        const TcStep& ts = *cur;

        if (warp == 8) {
            // ===== weight producer: stream this step's chunks through the ring (one elected lane) =====
            for (int ci = 0; ci < ts.nchunks; ++ci, ++prod_it) {
                if ((tid & 31) == 0) {
                    const uint32_t slot = prod_it % kTcSlots;
                    mbar_wait(&empty[slot], ((prod_it / kTcSlots) & 1u) ^ 1u);
                    mbar_expect_tx(&full[slot], (uint32_t)ts.chunk_bytes[ci]);
                    bulk_load(slots + (size_t)slot * dm.slot_bytes, img + ts.chunk_off[ci], (uint32_t)ts.chunk_bytes[ci], &full[slot]);
                }
                __syncwarp();
            }
            continue;
        }

        // ===== compute group: thread = one row of the tile =====
        const int C = ts.C, c = ts.c, Kin = ts.Kin, Ku = ts.Ku, HPd = ts.HPd, CPd = ts.CPd, Hc = ts.Hc;
        const long long row0 = (pair * 2 + x.g) * 128;
        long long n = row0 + x.tid;
        const bool valid = n < a.N;
        if (!valid) n = a.N - 1;
        const long long ag = n / a.PPA, rem = n % a.PPA, pt = rem / a.K;
        float z0, z1, prev0, prev1, lacc = 0.0f;
        {
            const float* xp = src_at(a.x, ag, pt, step);
            z0 = __ldg(xp); z1 = __ldg(xp + 1);
            const float* pp = step == 0 ? src_at(a.base, ag, pt, 0) : src_at(a.prefix, ag, pt, step - 1);
            prev0 = __ldg(pp); prev1 = __ldg(pp + 1);
        }
        // ---- input operand [z (2), cond (C0), prefix (2*step), 0...] ----
        {
            const float* cp = src_at(a.cond, ag, pt, step);
            for (int k0 = 0; k0 < Kin; k0 += 8) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int col = k0 + i;
                    float val = 0.0f;
                    if (col < 2) val = col == 0 ? z0 : z1;
                    else if (col < 2 + dm.C0) val = __ldg(cp + (col - 2));
                    else if (col < c) { const int e = col - 2 - dm.C0; val = __ldg(src_at(a.prefix, ag, pt, e >> 1) + (e & 1)); }
                    v[i] = val;
                }
                *reinterpret_cast<uint4*>(x.A_in + kmajor_offset(x.tid, k0, Kin)) =
                    make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
            }
        }
        fence_async_smem();
        tc_fence_before();
        group_bar(x);

        float s_ls_r = 0.0f, s_e2 = 0.0f;
        // ================= density nets: which = 0 (r, sample) then couplings then which = 1 (q, eval) =======
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            // ---- L1 ----
            {
                const uint8_t* ch = chunk_wait(x);
                if (x.tid == 0) {
                    tc_fence_after();
                    umma_tile(x.tacc, smem_u32(x.A_in), smem_u32(ch), Kin, 2 * HPd, false);
                    umma_commit(x.mma_bar);
                }
                mma_wait(x);
                float d0, d1;
                epi_hidden<A_RELU, A_RELU, false>(x, HPd, reinterpret_cast<const float*>(ch + (size_t)2 * HPd * Kin * 2), nullptr, d0, d1);
                layer_end(x);
            }
            // ---- L2 ----
            {
                const uint8_t* ch = chunk_wait(x);
                if (x.tid == 0) {
                    tc_fence_after();
                    umma_tile(x.tacc, smem_u32(x.A_h0), smem_u32(ch), HPd, HPd, false);
                    umma_tile(x.tacc + HPd, smem_u32(x.A_h1), smem_u32(ch + (size_t)HPd * HPd * 2), HPd, HPd, false);
                    umma_commit(x.mma_bar);
                }
                mma_wait(x);
                float d0, d1;
                epi_hidden<A_RELU, A_RELU, false>(x, HPd, reinterpret_cast<const float*>(ch + (size_t)2 * HPd * HPd * 2), nullptr, d0, d1);
                layer_end(x);
            }
            // ---- L3 -> means [0,CPd), log-stddevs [CPd, 2CPd) ----
            {
                const uint8_t* ch = chunk_wait(x);
                if (x.tid == 0) {
                    tc_fence_after();
                    umma_tile(x.tacc, smem_u32(x.A_h0), smem_u32(ch), HPd, CPd, false);
                    umma_tile(x.tacc + CPd, smem_u32(x.A_h1), smem_u32(ch + (size_t)CPd * HPd * 2), HPd, CPd, false);
                    umma_commit(x.mma_bar);
                }
                mma_wait(x);
                const float* bias = reinterpret_cast<const float*>(ch + (size_t)2 * CPd * HPd * 2);
                if (which == 0) {
                    // sample u = exp(ls) eps + mu (fastpredNF.py:727-733); log r = -C/2 log 2pi - sum ls - 1/2 sum eps^2
                    const int m0 = ts.masked_dim[0];
                    for (int d0 = 0; d0 < Ku; d0 += 8) {
                        float u[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        if (d0 < CPd) {
                            float mu[8], ls[8], e[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                            tmem_ld8(x.tld + (uint32_t)d0, mu);
                            tmem_ld8(x.tld + (uint32_t)(CPd + d0), ls);
                            if (a.eps != nullptr) {
                                const float* ep = a.eps + n * a.eps_rs + ts.eps_off + d0;
#pragma unroll
                                for (int i = 0; i < 8; ++i) if (d0 + i < C) e[i] = __ldg(ep + i);
                            } else {
                                const float4 g0 = philox_normal4(a.seed, n, (unsigned)step, (unsigned)(d0 >> 2));
                                const float4 g1 = philox_normal4(a.seed, n, (unsigned)step, (unsigned)(d0 >> 2) + 1u);
                                e[0] = g0.x; e[1] = g0.y; e[2] = g0.z; e[3] = g0.w; e[4] = g1.x; e[5] = g1.y; e[6] = g1.z; e[7] = g1.w;
                            }
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                if (d0 + i < C) {
                                    const float m = mu[i] + bias[d0 + i], l = ls[i] + bias[CPd + d0 + i];
                                    u[i] = fmaf(__expf(l), e[i], m);
                                    s_ls_r += l;
                                    s_e2 = fmaf(e[i], e[i], s_e2);
                                }
                            }
                            tmem_st8(x.tld + (uint32_t)(kTcUCol + d0), u);
                        }
                        if (d0 == (C & ~7)) {        // mx = z * mask of the first coupling sits at columns C, C+1
                            u[(C & 7)] = m0 == 0 ? z0 : 0.0f;
                            u[(C & 7) + 1] = m0 == 1 ? z1 : 0.0f;
                        }
                        *reinterpret_cast<uint4*>(x.A_u + kmajor_offset(x.tid, d0, Ku)) =
                            make_uint4(pack_bf16x2(u[0], u[1]), pack_bf16x2(u[2], u[3]), pack_bf16x2(u[4], u[5]), pack_bf16x2(u[6], u[7]));
                    }
                    tmem_wait_st();
                    lacc -= (-0.5f * (float)C * kLog2Pi - s_ls_r) - 0.5f * s_e2;
                } else {
                    // log q(u | w, y) (fastpredNF.py:713-724) with u read back from TMEM
                    float s_ls = 0.0f, s_q = 0.0f;
                    for (int d0 = 0; d0 < CPd; d0 += 8) {
                        float mu[8], ls[8], u[8];
                        tmem_ld8(x.tld + (uint32_t)d0, mu);
                        tmem_ld8(x.tld + (uint32_t)(CPd + d0), ls);
                        tmem_ld8(x.tld + (uint32_t)(kTcUCol + d0), u);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (d0 + i < C) {
                                const float m = mu[i] + bias[d0 + i], l = ls[i] + bias[CPd + d0 + i];
                                const float df = (u[i] - m) * __expf(-l);
                                s_q = fmaf(df, df, s_q);
                                s_ls += l;
                            }
                        }
                    }
                    lacc += (-0.5f * (float)C * kLog2Pi - s_ls) - 0.5f * s_q;
                }
                layer_end(x);
            }
            if (which == 1) break;

            // ================= n_blocks x (masked affine coupling + BatchNorm), forward =================
#pragma unroll 1
            for (int b = 0; b < ts.n_blocks; ++b) {
                const int m = ts.masked_dim[b], j = 1 - m;
                float ds = 0.0f, dt = 0.0f;
                const float* tail = nullptr;
                {   // L1: [u, mx] -> 2H
                    const uint8_t* ch = chunk_wait(x);
                    if (x.tid == 0) {
                        tc_fence_after();
                        umma_tile(x.tacc, smem_u32(x.A_u), smem_u32(ch), Ku, 2 * Hc, false);
                        umma_commit(x.mma_bar);
                    }
                    mma_wait(x);
                    tail = reinterpret_cast<const float*>(ch + (size_t)2 * Hc * Ku * 2);
                    if (ts.n_hidden == 0) epi_hidden<A_TANH, A_RELU, true>(x, Hc, tail, tail + 2 * Hc, ds, dt);
                    else epi_hidden<A_TANH, A_RELU, false>(x, Hc, tail, nullptr, ds, dt);
                    if (ts.n_hidden > 0) layer_end(x);
                }
                for (int l = 1; l <= ts.n_hidden; ++l) {
                    const uint8_t* ch = chunk_wait(x);
                    if (x.tid == 0) {
                        tc_fence_after();
                        umma_tile(x.tacc, smem_u32(x.A_h0), smem_u32(ch), Hc, Hc, false);
                        umma_tile(x.tacc + Hc, smem_u32(x.A_h1), smem_u32(ch + (size_t)Hc * Hc * 2), Hc, Hc, false);
                        umma_commit(x.mma_bar);
                    }
                    mma_wait(x);
                    tail = reinterpret_cast<const float*>(ch + (size_t)2 * Hc * Hc * 2);
                    if (l == ts.n_hidden) epi_hidden<A_TANH, A_RELU, true>(x, Hc, tail, tail + 2 * Hc, ds, dt);
                    else { epi_hidden<A_TANH, A_RELU, false>(x, Hc, tail, nullptr, ds, dt); layer_end(x); }
                }
                // affine update of the unmasked coordinate + BatchNorm (TFCondARFlow.py:360-381, 303-315)
                {
                    const float s = ds + tail[4 * Hc], t = dt + tail[4 * Hc + 1];
                    const float log_s = tanhf(s);
                    float zz[2] = {z0, z1};
                    zz[j] = fmaf(zz[j], __expf(log_s), t);
                    z0 = fmaf(ts.bn_scale[b][0], zz[0], ts.bn_shift[b][0]);
                    z1 = fmaf(ts.bn_scale[b][1], zz[1], ts.bn_shift[b][1]);
                    lacc += log_s + ts.bn_ldj[b];
                    if (b + 1 < ts.n_blocks) {
                        const int mn = ts.masked_dim[b + 1];
                        *reinterpret_cast<uint32_t*>(x.A_u + kmajor_offset(x.tid, C, Ku)) =
                            pack_bf16x2(mn == 0 ? z0 : 0.0f, mn == 1 ? z1 : 0.0f);
                    } else {
                        *reinterpret_cast<uint32_t*>(x.A_in + kmajor_offset(x.tid, 0, Kin)) = pack_bf16x2(z0, z1);
                    }
                }
                layer_end(x);
            }
        }
        {     // whole warps (row = thread): store_result reduces K importance samples across lanes
            const float lp = normal_lp2(z0, z1, prev0, prev1, ts.std[0], ts.std[1]) + lacc;
            store_result(a, ag, rem, step, lp, valid);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 512);
}

inline int tc_launch_separate(const TcDims& dm, const TcStep* d_steps, const unsigned char* d_img, const RowArgs& a,
                              int num_sms, int smem_optin, cudaStream_t stream, uint64_t* launches, std::string* msg) {
    if (!dm.supported || d_img == nullptr) {
        *msg = "the bf16 tensor-core path does not support this model shape (needs hidden_size % 16 == 0, 2*hidden_size <= 192, "
               "4*(cond_size + 2*pred_len) <= 192); use FC_PREC_FP32";
        return 1;
    }
    const size_t smem = tc_smem_bytes(dm);
    if ((int)smem > smem_optin) { *msg = "tensor-core path: shared-memory budget exceeded"; return 1; }
    cudaError_t e = cudaFuncSetAttribute(k_separate_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { *msg = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return 1; }
    const long long tiles = (a.N + 127) / 128;
    const long long pairs = (tiles + 1) / 2;
    const long long items = pairs * dm.T;
    const int grid = (int)(items < num_sms ? items : num_sms);
    k_separate_tc<<<grid, kTcThreads, smem, stream>>>(d_steps, d_img, a, dm, pairs, items);
    (*launches)++;
    e = cudaGetLastError();
    if (e != cudaSuccess) { *msg = std::string("k_separate_tc launch: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

}  // namespace fc
