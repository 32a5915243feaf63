// kernels_tc2.cuh -- accurate tensor-core path of the fused CIF step ("f16x3"):
//   * every Linear runs on tcgen05 with fp16 operands and fp32 accumulation; with NPASS = 3 operands are split
//     a = a_hi + a_lo, W = W_hi + W_lo and D = a_hi W_hi + a_lo W_hi + a_hi W_lo (~22 mantissa bits, fp32-grade);
//     NPASS = 1 is the single-pass fp16 mode (11 mantissa bits, ~3x fewer MMAs);
//   * activations never leave the SM: accumulators AND the next layer's A operand live in tensor memory
//     (tcgen05.st of packed fp16 pairs, TS-mode MMA), shared memory only holds the weight ring fed by TMA bulk
//     copies, so one CTA keeps a 128-row tile plus a deep weight prefetch resident;
//   * the two independent nets of every sub-network (s|t of a coupling, shift|log-scale of r/q) are interleaved:
//     while the CUDA cores run the epilogue of one net the tensor core runs the next layer of the other.
// One CTA = one 128-row tile; 8 epilogue warps (row = TMEM lane, two warps per lane quadrant split the columns)
// + 1 producer warp.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstring>
#include <string>
#include <vector>

#include "../../include/flowchain_b200.h"
#include "kernels_tc.cuh"
#include "plan.h"
#include "tc_common.cuh"

namespace fc {

constexpr int kT2MaxChunks = 12 + 2 * (kMaxHidden + 1) * kMaxBlocks;
constexpr uint32_t kT2Slots = 8;       // weight-ring depth (power of two: ring arithmetic is a mask/shift)
// tensor-memory column map (32-bit columns, 128 lanes = rows)
constexpr uint32_t kT2Acc0 = 0, kT2Acc1 = 96;            // fp32 accumulators of net 0 / net 1 (N <= 96)
constexpr uint32_t kT2InHi = 192, kT2InLo = 216;         // [z|w, cond, prefix] operand, K <= 48
constexpr uint32_t kT2UHi = 240, kT2ULo = 264;           // [u, mx] operand, K <= 48
constexpr uint32_t kT2H0Hi = 288, kT2H0Lo = 336;         // hidden operand of net 0, K <= 96
constexpr uint32_t kT2H1Hi = 384, kT2H1Lo = 432;         // hidden operand of net 1

struct T2Step {
    int32_t C, c, Kin, Ku, Wd, CPd, Hc;
    int32_t n_blocks, n_hidden, nchunks, eps_off, seq_eps_off;   // noise-tape offsets: per-step / chain ("sequential") layout
    int32_t chunk_off[kT2MaxChunks];
    int32_t chunk_bytes[kT2MaxChunks];
    int32_t masked_dim[kMaxBlocks];
    float bn_scale[kMaxBlocks][2], bn_shift[kMaxBlocks][2], bn_ldj[kMaxBlocks];
    float bn_iscale[kMaxBlocks][2], bn_ishift[kMaxBlocks][2];     // BatchNorm.inverse: x = iscale * y + ishift (sampler)
    float std[2];
    // MMA-issuer schedule, one entry per chunk (= net-layer), executed in order:
    //   wait on barrier `wait` (0 none, 1 opnd[0], 2 opnd[1], 3 fin) -> issue the MMAs -> hand back `rel` older chunks
    uint8_t job_wait[kT2MaxChunks], job_rel[kT2MaxChunks], job_net[kT2MaxChunks], job_src[kT2MaxChunks];  // src: 0 IN, 1 U, 2 H[net]
    uint8_t job_kp[kT2MaxChunks], job_n[kT2MaxChunks];      // Kp, N in elements
    uint8_t tail_rel, pad1[3];                                // chunks handed back after the last `fin`
    // 1: the chunk ends with a [N x 16] fp16 bias tile (columns 0..2 = the bias as three fp16 pieces) that the two-tile
    // kernel multiplies by a constant ones tile, so the bias arrives inside the accumulator (kernels_tc3.cuh)
    uint8_t job_fold[kT2MaxChunks];
};

struct T2Dims {
    int32_t slot_bytes, nslots;
    int32_t T, C0;
    int32_t supported;
    int32_t step_hi;         // a launch covers steps step_hi, step_hi-1, ... (heavy first); set by the launcher
};


constexpr float kTanhPrescale = 2.885390081777927f;     // 2 log2(e)

// hidden layers up to this width carry a bias tile (wider ones would not fit the two-tile kernel's weight slots)
constexpr int kFoldMaxN = 64;

// Packs the fp16 image: per step, per net-layer one chunk = [hi tile][lo tile (npass 3)][fp32 tail][bias tile (folded)].
inline void tc2_pack(const fc_model_desc& d, const std::vector<StepPlan>& plans, const std::vector<TcStepSrc>& src, int npass,
                     std::vector<unsigned char>& img, std::vector<T2Step>& steps, T2Dims& dm, int smem_optin) {
    img.clear();
    steps.clear();
    memset(&dm, 0, sizeof dm);
    const int T = d.pred_len, H = d.hidden_size;
    const int cmax = d.cond_size + (T - 1) * 2 + 2;
    if (H % 16 != 0 || H > 96 || pad16i(2 * cmax) > 96 || pad16i(cmax) > 48) return;
    dm.T = T; dm.C0 = d.cond_size;
    steps.resize(T);
    std::vector<float> tmp;
    for (int i = 0; i < T; ++i) {
        T2Step& ts = steps[i];
        memset(&ts, 0, sizeof ts);
        const StepPlan& sp = plans[i];
        const TcStepSrc& s = src[i];
        const int C = sp.C, c = sp.c;
        ts.C = C; ts.c = c; ts.Kin = pad16i(c); ts.Ku = pad16i(C + 2); ts.Wd = pad16i(2 * c); ts.CPd = pad16i(C); ts.Hc = H;
        ts.n_blocks = sp.n_blocks; ts.n_hidden = sp.n_hidden; ts.eps_off = sp.eps_off; ts.seq_eps_off = sp.seq_eps_off;
        for (int b = 0; b < sp.n_blocks; ++b) {
            ts.masked_dim[b] = sp.masked_dim[b];
            ts.bn_scale[b][0] = sp.bn_scale[b][0]; ts.bn_scale[b][1] = sp.bn_scale[b][1];
            ts.bn_shift[b][0] = sp.bn_shift[b][0]; ts.bn_shift[b][1] = sp.bn_shift[b][1];
            ts.bn_ldj[b] = sp.bn_ldj[b];
            ts.bn_iscale[b][0] = sp.bn_iscale[b][0]; ts.bn_iscale[b][1] = sp.bn_iscale[b][1];
            ts.bn_ishift[b][0] = sp.bn_ishift[b][0]; ts.bn_ishift[b][1] = sp.bn_ishift[b][1];
        }
        ts.std[0] = sp.std[0]; ts.std[1] = sp.std[1];
        int nch = 0;
        // one chunk: W [rows x K] (row-major, as nn.Linear.weight) -> canonical K-major tiles + tail
        auto chunk = [&](const float* W, const float* bias, int rows, int K, int rows_pad, int Kp, const float* extra, int n_extra, bool fold) {
            while (img.size() % 128) img.push_back(0);
            ts.chunk_off[nch] = (int32_t)img.size();
            for (int part = 0; part < (npass == 3 ? 2 : 1); ++part) {
                const size_t o = img.size();
                img.resize(o + (size_t)rows_pad * Kp * 2);
                pack_kmajor_f16(img.data() + o, W, rows, K, rows_pad, Kp, part);
            }
            const size_t o = img.size();
            img.resize(o + (size_t)(rows_pad + n_extra) * 4, 0);
            memcpy(img.data() + o, bias, (size_t)rows * 4);
            if (n_extra) memcpy(img.data() + o + (size_t)rows_pad * 4, extra, (size_t)n_extra * 4);
            while (img.size() % 16) img.push_back(0);
            fold = fold && rows_pad <= kFoldMaxN;
            if (fold) {      // bias = p0 + p1 + p2 with fp16 pieces (33 mantissa bits): exact through the fp32 accumulate
                std::vector<float> bm((size_t)rows_pad * 16, 0.0f);
                for (int r = 0; r < rows; ++r) {
                    const float p0 = h2f(f2h(bias[r])), p1 = h2f(f2h(bias[r] - p0)), p2 = h2f(f2h(bias[r] - p0 - p1));
                    bm[(size_t)r * 16] = p0; bm[(size_t)r * 16 + 1] = p1; bm[(size_t)r * 16 + 2] = p2;
                }
                const size_t ob = img.size();
                img.resize(ob + (size_t)rows_pad * 16 * 2);
                pack_kmajor_f16(img.data() + ob, bm.data(), rows_pad, 16, rows_pad, 16, 0);
            }
            ts.job_fold[nch] = fold ? 1 : 0;
            ts.chunk_bytes[nch] = (int32_t)(img.size() - ts.chunk_off[nch]);
            dm.slot_bytes = ts.chunk_bytes[nch] > dm.slot_bytes ? ts.chunk_bytes[nch] : dm.slot_bytes;
            ++nch;
        };
        auto dens = [&](int which) {
            const int h2 = 2 * c;
            const int dims[4] = {c, h2, h2, C};
            const int kp[3] = {ts.Kin, ts.Wd, ts.Wd}, np[3] = {ts.Wd, ts.Wd, ts.CPd};
            for (int l = 0; l < 3; ++l)
                for (int net = 0; net < 2; ++net)
                    chunk(s.dens_W[which][net][l], s.dens_b[which][net][l], dims[l + 1], dims[l], np[l], kp[l], nullptr, 0, l < 2);
        };
        dens(0);
        for (int b = 0; b < sp.n_blocks; ++b) {
            const int j = 1 - sp.masked_dim[b];
            for (int l = 0; l <= sp.n_hidden; ++l)
                for (int net = 0; net < 2; ++net) {
                    std::vector<float> extra;
                    if (l == sp.n_hidden) {      // surviving row (1 - mask) of the final H -> 2 Linear + its bias
                        extra.assign(H + 4, 0.0f);
                        memcpy(extra.data(), s.cpl_W[b][net][sp.n_hidden + 1] + (size_t)j * H, (size_t)H * 4);
                        extra[H] = s.cpl_b[b][net][sp.n_hidden + 1][j];
                    }
                    const float* Wl = s.cpl_W[b][net][l];
                    const float* bl = s.cpl_b[b][net][l];
                    const int Kl = l == 0 ? c : H;
                    std::vector<float> Ws, bs;
                    if (npass == 3 && net == 0) {   // s-net layers feed tanh = 1 - 2/(1 + 2^(2 log2(e) x)): fold the factor into W, b
                        Ws.assign(Wl, Wl + (size_t)H * Kl);
                        bs.assign(bl, bl + H);
                        for (float& v : Ws) v *= kTanhPrescale;
                        for (float& v : bs) v *= kTanhPrescale;
                        Wl = Ws.data(); bl = bs.data();
                    }
                    chunk(Wl, bl, H, Kl, H, l == 0 ? ts.Ku : H, extra.empty() ? nullptr : extra.data(), (int)extra.size(), true);
                }
        }
        dens(1);
        ts.nchunks = nch;
        {   // issuer schedule (see k_separate_tc2: compute warps arrive on fin / opnd[net] in exactly this order)
            int j = 0;
            auto put = [&](int wait, int net, int src_, int kp, int n, int rel) {
                ts.job_wait[j] = (uint8_t)wait; ts.job_net[j] = (uint8_t)net; ts.job_src[j] = (uint8_t)src_;
                ts.job_kp[j] = (uint8_t)kp; ts.job_n[j] = (uint8_t)n; ts.job_rel[j] = (uint8_t)rel; ++j;
            };
            auto density = [&](bool first) {
                // L1 of both nets reads the input operand; `fin` was awaited by the caller's first entry
                put(3, 0, 0, ts.Kin, ts.Wd, 0);
                put(3, 1, 0, ts.Kin, ts.Wd, 0);
                for (int l = 0; l < 2; ++l)
                    for (int net = 0; net < 2; ++net) put(1 + net, net, 2, ts.Wd, l == 0 ? ts.Wd : ts.CPd, 1);
            };
            density(true);
            for (int b = 0; b < sp.n_blocks; ++b) {
                put(3, 0, 1, ts.Ku, H, 0);
                put(3, 1, 1, ts.Ku, H, 0);
                for (int l = 0; l < sp.n_hidden; ++l)
                    for (int net = 0; net < 2; ++net) put(1 + net, net, 2, H, H, 1);
            }
            density(false);
            ts.tail_rel = 2;                                          // the q density's two L3 chunks (after the last fin)
        }
    }
    while (img.size() % 128) img.push_back(0);
    dm.slot_bytes = (dm.slot_bytes + 127) / 128 * 128;
    dm.nslots = (int32_t)kT2Slots;
    dm.supported = 10240 + (size_t)kT2Slots * dm.slot_bytes <= (size_t)smem_optin;
}

#ifndef FC_T2_PROF
#define FC_T2_PROF 0
#endif
__device__ unsigned long long g_t2prof[16];
#define T2_TICK(i)                                                                      \
    if (FC_T2_PROF && blockIdx.x == 0 && threadIdx.x == 0) {                            \
        long long now_;                                                                 \
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(now_));                            \
        atomicAdd(&g_t2prof[i], (unsigned long long)(now_ - prof_t));                   \
        prof_t = now_;                                                                  \
    }

#ifndef FC_T2_TRACE
#define FC_T2_TRACE 0
#endif
// development aid: event timeline of one item (CTA 0) -- g_trace[warp][k] = (tag << 48) | clock
__device__ unsigned long long g_t2trace[12][256];
#define T2_EV(tag)                                                                                   \
    if (FC_T2_TRACE && blockIdx.x == 0 && trace_on && (threadIdx.x & 31) == 0 && trace_n < 256) {    \
        long long now_;                                                                              \
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(now_));                                         \
        g_t2trace[threadIdx.x >> 5][trace_n++] = ((unsigned long long)(tag) << 48) | ((unsigned long long)now_ & 0xFFFFFFFFFFFFull); \
    }

namespace t2 {
using namespace tc;

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r) : "memory");
}

__device__ __forceinline__ float clamp_h(float v) { return fminf(fmaxf(v, -65000.0f), 65000.0f); }

// 8 fp32 values -> 4 packed fp16 pairs (hi) and the 4 packed residual pairs (lo), with an optional fused ReLU.
// NPASS = 3: hi is the value TRUNCATED to fp16's 11 significant bits (one LOP3 on the integer pipe; exactly
// representable, so the pack is exact), lo = a - hi in fp32 (same sign as a), both packed with
// F2FP[.RELU][.SATFINITE] -- ReLU and the fp16 overflow clamp ride on the pack instruction for free.
// NPASS = 1: hi is the value rounded to fp16.
template <int NPASS, bool RELU>
__device__ __forceinline__ void split8(const float* a, uint32_t* hi, uint32_t* lo) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a0 = a[2 * i], a1 = a[2 * i + 1];
        if (NPASS == 3) {
            const float h0 = __uint_as_float(__float_as_uint(a0) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(a1) & 0xFFFFE000u);
            const float l0 = a0 - h0, l1 = a1 - h1;
            if (RELU) {
                asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi[i]) : "f"(h1), "f"(h0));
                asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(lo[i]) : "f"(l1), "f"(l0));
            } else {
                asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi[i]) : "f"(h1), "f"(h0));
                asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo[i]) : "f"(l1), "f"(l0));
            }
        } else {
            if (RELU) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi[i]) : "f"(a1), "f"(a0));
            else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi[i]) : "f"(a1), "f"(a0));
        }
    }
}
// tanh(x) = 1 - 2 / (1 + e^{2x}), |err| ~ 5e-7 (MUFU.TANH itself is only 2^-11 accurate).  Two values share ONE
// reciprocal (r = 1/(y0 y1), 1/y0 = r y1, 1/y1 = r y0): 1.5 MUFU ops per value instead of 2.
// The inputs arrive PRE-SCALED by 2 log2(e) (tc2_pack folds the factor into the s-net weights and biases).
__device__ __forceinline__ void tanh_acc2(float& x0, float& x1) {
    float e0, e1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(x0, 43.0f)));   // clamp: keeps y0*y1 finite, tanh(43 / 2.885) == 1
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(x1, 43.0f)));
    const float y0 = e0 + 1.0f, y1 = e1 + 1.0f;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y0 * y1));
    r *= -2.0f;
    x0 = fmaf(y1, r, 1.0f);
    x1 = fmaf(y0, r, 1.0f);
}
__device__ __forceinline__ float tanh_acc(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
    return 1.0f - __fdividef(2.0f, e + 1.0f);
}

// Barrier block in shared memory
struct Bars {
    uint64_t full[12], empty[12];    // weight ring: producer <-> MMA warp
    uint64_t mma_done[2];            // tcgen05.commit of the newest layer of net 0 / 1
    uint64_t opnd[2];                // all compute warps finished the epilogue of net 0 / 1 (its next operand is in TMEM)
    uint64_t fin;                    // all compute warps finished a sub-network (input operand of the next one is ready)
};

struct X {
    int row, part, lane;        // TMEM lane (tile row), column part id in [0, NPARTS), lane in warp
    uint32_t tb;                // tensor-memory base with this warp's lane quadrant
    Bars* bars;
    uint8_t* slots;
    int slot_bytes, nslots;
    uint32_t chunk_it;          // ring position of the chunk whose epilogue runs next (CTA lifetime)
    uint32_t issue_it, rel_it;  // ring positions of the next chunk to issue / to hand back (tracked by every warp)
    uint32_t tb0;               // tensor-memory base, lane 0 (MMA operands)
    uint32_t mma_cnt0, mma_cnt1;
    float* part_buf;            // [4][NPARTS][128] row-partial exchange
};

__device__ __forceinline__ const uint8_t* slot_ptr(const X& x, uint32_t it) { return x.slots + (size_t)(it % kT2Slots) * x.slot_bytes; }
__device__ __forceinline__ void wait_full(const X& x, uint32_t it) { mbar_wait(&x.bars->full[it % kT2Slots], (it / kT2Slots) & 1u); }
__device__ __forceinline__ void wait_mma(X& x, int net) {
    if (net == 0) { mbar_wait(&x.bars->mma_done[0], x.mma_cnt0 & 1u); x.mma_cnt0++; }
    else { mbar_wait(&x.bars->mma_done[1], x.mma_cnt1 & 1u); x.mma_cnt1++; }
    tc_fence_after();
}
// a compute warp tells the MMA warp that its tensor-memory reads/writes for this stage are complete
__device__ __forceinline__ void warp_arrive(const X& x, uint64_t* bar) {
    tmem_wait_ld();
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (x.lane == 0) mbar_arrive(bar);
}

enum { ACT_R = 0, ACT_T = 1 };

__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}

// hidden-layer epilogue of one net: acc + bias -> act -> (hi, lo) fp16 operand of the next layer; with DOT the fp32
// activations are contracted with the surviving row of the final Linear (returns this thread's partial).
// The 8-column chunks of the layer are dealt round-robin to the NPARTS warps that share a lane quadrant; a warp
// issues ALL its tensor-memory loads, waits once, then works through the values with full ILP.
// bias / wlast are shared-memory addresses (32-bit) into the weight ring.
template <int NPASS, int ACT, bool DOT, bool WRITE>
__device__ __forceinline__ void epi_chunk(float* u, int c0, uint32_t bias, uint32_t wlast, uint32_t ohi, uint32_t olo, float& dot) {
    const float4 b0 = lds128(bias + (uint32_t)c0 * 4), b1 = lds128(bias + (uint32_t)c0 * 4 + 16);
    u[0] += b0.x; u[1] += b0.y; u[2] += b0.z; u[3] += b0.w; u[4] += b1.x; u[5] += b1.y; u[6] += b1.z; u[7] += b1.w;
    if (ACT == ACT_T) {
#pragma unroll
        for (int q = 0; q < 8; q += 2) {
            if (NPASS == 3) tanh_acc2(u[q], u[q + 1]);
            else { asm("tanh.approx.f32 %0, %0;" : "+f"(u[q])); asm("tanh.approx.f32 %0, %0;" : "+f"(u[q + 1])); }
        }
    } else if (DOT) {
#pragma unroll
        for (int q = 0; q < 8; ++q) u[q] = fmaxf(u[q], 0.0f);      // (with WRITE the ReLU is fused into the fp16 pack)
    }
    if (DOT) {
        const float4 w0 = lds128(wlast + (uint32_t)c0 * 4), w1 = lds128(wlast + (uint32_t)c0 * 4 + 16);
        dot = fmaf(u[0], w0.x, dot); dot = fmaf(u[1], w0.y, dot); dot = fmaf(u[2], w0.z, dot); dot = fmaf(u[3], w0.w, dot);
        dot = fmaf(u[4], w1.x, dot); dot = fmaf(u[5], w1.y, dot); dot = fmaf(u[6], w1.z, dot); dot = fmaf(u[7], w1.w, dot);
    }
    if (WRITE) {
        uint32_t hi[4], lo[4];
        split8<NPASS, ACT == ACT_R>(u, hi, lo);
        tmem_st4(ohi + (uint32_t)(c0 >> 1), hi);
        if (NPASS == 3) tmem_st4(olo + (uint32_t)(c0 >> 1), lo);
    }
}

// One instance per (activation, role) -- deliberately NOT inlined: the kernel has ~10 epilogue sites and the
// instruction cache (32 KB) must hold the whole steady-state loop.  The warp owns 8-column chunks part, part+NPARTS,
// ...; two chunks are loaded per tensor-memory wait so 16 values are in flight per thread.
template <int NPASS, int NPARTS, int ACT, bool DOT, bool WRITE>
__device__ __noinline__ float epi_hidden_impl(uint32_t acc, uint32_t ohi, uint32_t olo, uint32_t bias, uint32_t wlast, int part, int W) {
    float dot = 0.0f;
#pragma unroll 1
    for (int c0 = part * 8; c0 < W; c0 += 16 * NPARTS) {
        const int c1 = c0 + 8 * NPARTS;
        float v0[8], v1[8];
        tmem_ld8(acc + (uint32_t)c0, v0);
        if (c1 < W) tmem_ld8(acc + (uint32_t)c1, v1);
        tmem_wait_ld();
        epi_chunk<NPASS, ACT, DOT, WRITE>(v0, c0, bias, wlast, ohi, olo, dot);
        if (c1 < W) epi_chunk<NPASS, ACT, DOT, WRITE>(v1, c1, bias, wlast, ohi, olo, dot);
    }
    return dot;
}

#ifndef FC_T2_ABLATE
#define FC_T2_ABLATE 0
#endif
template <int NPASS, int NPARTS, int ACT, bool DOT, bool WRITE>
__device__ __forceinline__ float epi_hidden(const X& x, int net, int W, uint32_t bias, uint32_t wlast) {
    if (FC_T2_ABLATE & 1) return 0.0f;
    return epi_hidden_impl<NPASS, NPARTS, ACT, DOT, WRITE>(x.tb + (net ? kT2Acc1 : kT2Acc0), x.tb + (net ? kT2H1Hi : kT2H0Hi),
                                                            x.tb + (net ? kT2H1Lo : kT2H0Lo), bias, wlast, x.part, W);
}

// sum of a per-row partial over the NPARTS column parts (every thread of the row gets the total)
template <int NPARTS>
__device__ __forceinline__ float row_sum(const X& x, int slot, float v) {
    x.part_buf[(slot * NPARTS + x.part) * 128 + x.row] = v;
    asm volatile("bar.sync 1, %0;" ::"n"(NPARTS * 128) : "memory");
    float s = 0.0f;
#pragma unroll
    for (int p = 0; p < NPARTS; ++p) s += x.part_buf[(slot * NPARTS + p) * 128 + x.row];
    return s;
}

// MMA warp (all lanes, convergent; one elected lane issues): acc[net][128 x N] = A[128 x Kp] * W[N x Kp]^T with 1 or 3 passes over the fp16 operand parts
template <int NPASS>
__device__ __noinline__ void mma_issue(Bars* bars, const uint8_t* chunk, uint32_t tb0, int net, uint32_t a_hi, uint32_t a_lo, int Kp, int N) {
    const uint32_t b_hi = smem_u32(chunk), b_lo = b_hi + (uint32_t)(N * Kp * 2);
    const uint32_t idesc = umma_idesc_f16(128, N), sbo = (uint32_t)(Kp >> 3) * 128u;
    const uint32_t acc = tb0 + (net ? kT2Acc1 : kT2Acc0);
    const uint64_t dhi = umma_desc(b_hi, 128, sbo), dlo = umma_desc(b_lo, 128, sbo);
    const int nk = (FC_T2_ABLATE & 2) ? 0 : (Kp >> 4);
    // The per-MMA operand set-up (address adds + moves into uniform registers) has ~100 cycles of dependent latency;
    // unrolled by 6 (= the largest K/16) the set-ups of one pass overlap and the loop is paced by the tensor pipe.
    uint32_t leader;
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(leader));
    if (leader) {
#pragma unroll 6
        for (int ks = 0; ks < nk; ++ks) umma_ts(acc, tb0 + a_hi + ks * 8, dhi + (uint64_t)(ks * 16), idesc, ks > 0 ? 1u : 0u);
        if (NPASS == 3) {
#pragma unroll 6
            for (int ks = 0; ks < nk; ++ks) umma_ts(acc, tb0 + a_lo + ks * 8, dhi + (uint64_t)(ks * 16), idesc, 1u);
#pragma unroll 6
            for (int ks = 0; ks < nk; ++ks) umma_ts(acc, tb0 + a_hi + ks * 8, dlo + (uint64_t)(ks * 16), idesc, 1u);
        }
        umma_commit(&bars->mma_done[net]);
    }
    __syncwarp();
}

}  // namespace t2

// Warp roles: [0, 4*NPARTS) compute (warp w: lane quadrant w % 4, column part w / 4), then the weight producer (TMA)
// and two MMA issuers (one per net).  No CTA-wide barrier inside an item: roles hand work to each other through mbarriers only
// (epilogue(l) -> `opnd` -> MMA(l+1) -> tcgen05.commit -> `mma_done` -> epilogue(l+1)).
template <int NPASS, int NPARTS>
__global__ void __launch_bounds__(NPARTS * 128 + 96, 1)
k_separate_tc2(const T2Step* __restrict__ steps, const uint8_t* __restrict__ img, RowArgs a, T2Dims dm, long long tiles_per_step,
               long long total_items) {
    using namespace tc;
    using namespace t2;
    constexpr int NCW = 4 * NPARTS;                  // compute warps
    constexpr int kThreadsAll = NPARTS * 128 + 96;
    extern __shared__ __align__(128) uint8_t smem_t2[];
    uint8_t* smem = smem_t2;
    Bars* bars = reinterpret_cast<Bars*>(smem);
    static_assert(sizeof(Bars) <= 240, "barrier block");
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + 240);
    T2Step* cur = reinterpret_cast<T2Step*>(smem + 256);
    static_assert(sizeof(T2Step) <= 1792, "T2Step must fit the header");
    float* part = reinterpret_cast<float*>(smem + 2048);         // 4 x NPARTS x 128 floats (<= 8 KB) -> header is 10240 B
    uint8_t* slots = smem + 10240;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < (int)kT2Slots; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->mma_done[0], 1);
        mbar_init(&bars->mma_done[1], 1);
        mbar_init(&bars->opnd[0], NCW);
        mbar_init(&bars->opnd[1], NCW);
        mbar_init(&bars->fin, NCW);
        mbar_fence_init();
    }
    if (warp == NCW) tmem_alloc(tslot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tslot;

    X x;
    x.lane = lane; x.part = warp >> 2; x.row = (warp & 3) * 32 + lane;
    x.tb0 = __shfl_sync(0xffffffffu, tmem_base, 0);
    x.tb = x.tb0 + ((uint32_t)((warp & 3) * 32) << 16);
    x.bars = bars; x.slots = slots; x.slot_bytes = dm.slot_bytes; x.nslots = dm.nslots;
    x.chunk_it = 0; x.issue_it = 0; x.rel_it = 0; x.mma_cnt0 = 0; x.mma_cnt1 = 0; x.part_buf = part;
    uint32_t ring_it = 0;            // producer / MMA-warp ring position
    uint32_t opnd_cnt0 = 0, opnd_cnt1 = 0, fin_cnt = 0;

    long long prof_t = 0;
    if (FC_T2_PROF) asm volatile("mov.u64 %0, %%clock64;" : "=l"(prof_t));
    int trace_n = 0, item_no = 0;
    bool trace_on = false;
    int cur_step = -1;
    for (long long item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int step = dm.step_hi - (int)(item / tiles_per_step);    // heavy (late) steps first
        const long long tile = item % tiles_per_step;
        trace_on = (item_no++ == 5);
        T2_EV(1)
        T2_TICK(0)   // tail of previous item (output store)
        if (step != cur_step) {              // (CTA-uniform) consecutive items of a CTA almost always share the step
            __syncthreads();
            const int nw = sizeof(T2Step) / 4;
            const int* s = reinterpret_cast<const int*>(steps + step);
            int* d = reinterpret_cast<int*>(cur);
            for (int i = tid; i < nw; i += kThreadsAll) d[i] = __ldg(s + i);
            cur_step = step;
            __syncthreads();
        }
        const T2Step& ts = *cur;
        const int C = ts.C, c = ts.c, Kin = ts.Kin, Ku = ts.Ku, Wd = ts.Wd, CPd = ts.CPd, Hc = ts.Hc;

        if (warp == NCW) {
            // ===== weight producer: stream the step's chunks through the ring =====
            for (int ci = 0; ci < ts.nchunks; ++ci, ++ring_it) {
                if (lane == 0) {
                    const uint32_t slot = ring_it % kT2Slots;
                    mbar_wait(&bars->empty[slot], ((ring_it / kT2Slots) & 1u) ^ 1u);
                    if (FC_T2_ABLATE & 4) mbar_arrive(&bars->full[slot]);
                    else {
                    mbar_expect_tx(&bars->full[slot], (uint32_t)ts.chunk_bytes[ci]);
                    bulk_load(slots + (size_t)slot * dm.slot_bytes, img + ts.chunk_off[ci], (uint32_t)ts.chunk_bytes[ci], &bars->full[slot]);
                    }
                }
                __syncwarp();
            }
            continue;
        }

        if (warp > NCW) {
            // ===== MMA issuers, one warp per net (whole warp convergent, one elected lane issues).  Issuing a
            // net-layer costs ~1-2k cycles of this warp's time (operand set-up through uniform registers), so one
            // issuer for both nets was the critical path; with one per net the two chains are independent.
            // Net n owns chunks n, n+2, ... of the step; after issuing job j it hands back the chunk of job j-1,
            // whose epilogue finished before the barrier this job waited on. =====
            const int net = warp - NCW - 1;
            uint64_t* opnd_bar = &bars->opnd[net];
            uint64_t* done_bar = &bars->mma_done[net];
            const uint32_t base = ring_it;
            const uint32_t acc = x.tb0 + (net ? kT2Acc1 : kT2Acc0);
            int prev = -1;
#pragma unroll 1
            for (int ci = net; ci < ts.nchunks; ci += 2) {
                // ---- everything that does not depend on the operand being ready: weights present, descriptors ----
                const uint32_t it = base + (uint32_t)ci;
                mbar_wait(&bars->full[it % kT2Slots], (it / kT2Slots) & 1u);
                const int w = ts.job_wait[ci], src = ts.job_src[ci], Kp = ts.job_kp[ci], N = ts.job_n[ci];
                const uint32_t a_hi = x.tb0 + (src == 0 ? kT2InHi : (src == 1 ? kT2UHi : (net ? kT2H1Hi : kT2H0Hi)));
                const uint32_t a_lo = x.tb0 + (src == 0 ? kT2InLo : (src == 1 ? kT2ULo : (net ? kT2H1Lo : kT2H0Lo)));
                const uint32_t b_hi = smem_u32(slots + (size_t)(it % kT2Slots) * dm.slot_bytes);
                const uint32_t idesc = umma_idesc_f16(128, N), sbo = (uint32_t)(Kp >> 3) * 128u;
                const uint64_t dhi = umma_desc(b_hi, 128, sbo), dlo = umma_desc(b_hi + (uint32_t)(N * Kp * 2), 128, sbo);
                const int nk = (FC_T2_ABLATE & 2) ? 0 : (Kp >> 4);
                uint32_t leader;
                asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(leader));
                // ---- critical path: operand ready -> MMAs -> commit ----
                if (w == 3) { mbar_wait(&bars->fin, fin_cnt & 1u); fin_cnt++; }
                else { mbar_wait(opnd_bar, opnd_cnt0 & 1u); opnd_cnt0++; }
                tc_fence_after();
                T2_EV(300 + ci)
                if (leader) {
#pragma unroll 6
                    for (int ks = 0; ks < nk; ++ks) umma_ts(acc, a_hi + ks * 8, dhi + (uint64_t)(ks * 16), idesc, ks > 0 ? 1u : 0u);
                    if (NPASS == 3) {
#pragma unroll 6
                        for (int ks = 0; ks < nk; ++ks) umma_ts(acc, a_lo + ks * 8, dhi + (uint64_t)(ks * 16), idesc, 1u);
#pragma unroll 6
                        for (int ks = 0; ks < nk; ++ks) umma_ts(acc, a_hi + ks * 8, dlo + (uint64_t)(ks * 16), idesc, 1u);
                    }
                    umma_commit(done_bar);
                }
                __syncwarp();
                T2_EV(400 + ci)
                if (prev >= 0) mbar_arrive_elect(&bars->empty[(base + (uint32_t)prev) % kT2Slots]);
                prev = ci;
            }
            mbar_wait(&bars->fin, fin_cnt & 1u);                     // final q-density epilogue done
            fin_cnt++;
            mbar_arrive_elect(&bars->empty[(base + (uint32_t)prev) % kT2Slots]);
            ring_it += (uint32_t)ts.nchunks;
            continue;
        }

        // ===== compute warps: thread = (row, column part) =====
        T2_TICK(1)   // item prologue: syncs + step table
        long long n = tile * 128 + x.row;
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
        // ---- input operand [z (2), cond (C0), prefix (2*step), 0...]: 8-column chunks dealt over the parts ----
        {
            const float* cp = src_at(a.cond, ag, pt, step) - 2;
            const float* pf = src_at(a.prefix, ag, pt, 0);
            const long long pst = a.prefix.st;
            const int c0e = 2 + dm.C0;
            for (int k0 = x.part * 8; k0 < Kin; k0 += 8 * NPARTS) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int col = k0 + i;
                    float val = 0.0f;
                    if (col < 2) val = col == 0 ? z0 : z1;
                    else if (col < c0e) val = __ldg(cp + col);
                    else if (col < c) { const int e = col - c0e; val = __ldg(pf + (long long)(e >> 1) * pst + (e & 1)); }
                    v[i] = clamp_h(val);
                }
                uint32_t hi[4], lo[4];
                split8<NPASS, false>(v, hi, lo);
                tmem_st4(x.tb + kT2InHi + (uint32_t)(k0 >> 1), hi);
                if (NPASS == 3) tmem_st4(x.tb + kT2InLo + (uint32_t)(k0 >> 1), lo);
            }
        }
        T2_TICK(2)   // row setup + input gather
        // first sub-network: both r-density nets read the input operand
        warp_arrive(x, &bars->fin);
        T2_EV(2)

#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            // ============ density nets (shift | log-scale): two relu layers, nets interleaved ============
#pragma unroll 1
            for (int l = 0; l < 2; ++l) {
#pragma unroll 1
                for (int net = 0; net < 2; ++net) {
                    const uint32_t bias = smem_u32(slot_ptr(x, x.chunk_it)) + (uint32_t)((NPASS == 3 ? 2 : 1) * Wd * (l == 0 ? Kin : Wd) * 2);
                    wait_mma(x, net);        // (the issuing warp acquired the weight chunk; the commit orders it for us)
                    T2_EV(10 + which * 100 + l * 2 + net)
                    epi_hidden<NPASS, NPARTS, ACT_R, false, true>(x, net, Wd, bias, 0u);
                    T2_EV(20 + which * 100 + l * 2 + net)
                    warp_arrive(x, &bars->opnd[net]);
                    x.chunk_it++;
                }
            }
            T2_TICK(3 + which * 4)   // density hidden layers (3: r, 7: q)
            // ---- L3 outputs: means in acc0[0,CPd), log-stddevs in acc1[0,CPd) ----
            const size_t w3 = (size_t)(NPASS == 3 ? 2 : 1) * CPd * Wd * 2;
            const float* bias_mu = reinterpret_cast<const float*>(slot_ptr(x, x.chunk_it) + w3);
            const float* bias_ls = reinterpret_cast<const float*>(slot_ptr(x, x.chunk_it + 1) + w3);
            x.chunk_it += 2;
            wait_mma(x, 0);
            wait_mma(x, 1);
            T2_EV(30 + which * 100)
            T2_TICK(4 + which * 4)   // wait for L3 MMAs (4: r, 8: q)
            if (which == 0) {
                // sample u = exp(ls) eps + mu (fastpredNF.py:727-733) -> [u, mx] operand; log r = -C/2 log 2pi - sum ls - 1/2 sum eps^2
                const int m0 = ts.masked_dim[0];
                float p_ls = 0.0f, p_e2 = 0.0f;
                for (int d0 = x.part * 8; d0 < Ku; d0 += 8 * NPARTS) {
                    float u[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) u[i] = 0.0f;
                    if (d0 < CPd) {
                        float mu[8], ls[8], e[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) e[i] = 0.0f;
                        tmem_ld8(x.tb + kT2Acc0 + (uint32_t)d0, mu);
                        tmem_ld8(x.tb + kT2Acc1 + (uint32_t)d0, ls);
                        if (a.eps != nullptr) {
                            const float* ep = a.eps + n * a.eps_rs + ts.eps_off + d0;
#pragma unroll
                            for (int i = 0; i < 8; ++i) if (d0 + i < C) e[i] = __ldg(ep + i);
                        } else if (!(FC_T2_ABLATE & 8)) {
                            const Normal8 g = philox_normal8(a.seed, n, (unsigned)step, (unsigned)(d0 >> 2));
                            e[0] = g.a.x; e[1] = g.a.y; e[2] = g.a.z; e[3] = g.a.w; e[4] = g.b.x; e[5] = g.b.y; e[6] = g.b.z; e[7] = g.b.w;
                        }
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (d0 + i < C) {
                                const float m = mu[i] + bias_mu[d0 + i], l = ls[i] + bias_ls[d0 + i];
                                u[i] = clamp_h(fmaf(__expf(l), e[i], m));
                                p_ls += l;
                                p_e2 = fmaf(e[i], e[i], p_e2);
                            }
                        }
                    }
                    if (d0 == (C & ~7)) {        // mx = z * mask of the first coupling sits at columns C, C+1
#pragma unroll
                        for (int i = 0; i < 8; i += 2)
                            if (i == (C & 7)) { u[i] = m0 == 0 ? clamp_h(z0) : 0.0f; u[i + 1] = m0 == 1 ? clamp_h(z1) : 0.0f; }
                    }
                    uint32_t hi[4], lo[4];
                    split8<3, false>(u, hi, lo);     // u is always kept as a hi+lo pair: log q needs it back at full precision
                    tmem_st4(x.tb + kT2UHi + (uint32_t)(d0 >> 1), hi);
                    tmem_st4(x.tb + kT2ULo + (uint32_t)(d0 >> 1), lo);
                }
                const float s_ls = row_sum<NPARTS>(x, 0, p_ls), s_e2 = row_sum<NPARTS>(x, 1, p_e2);
                lacc -= (-0.5f * (float)C * kLog2Pi - s_ls) - 0.5f * s_e2;
            } else {
                // log q(u | w, y) (fastpredNF.py:713-724), u reconstructed from its hi+lo operand
                float p_ls = 0.0f, p_q = 0.0f;
                for (int d0 = x.part * 8; d0 < CPd; d0 += 8 * NPARTS) {
                    float mu[8], ls[8];
                    uint32_t uh[4], ul[4];
                    tmem_ld8(x.tb + kT2Acc0 + (uint32_t)d0, mu);
                    tmem_ld8(x.tb + kT2Acc1 + (uint32_t)d0, ls);
                    tmem_ld4(x.tb + kT2UHi + (uint32_t)(d0 >> 1), uh);
                    tmem_ld4(x.tb + kT2ULo + (uint32_t)(d0 >> 1), ul);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (d0 + i < C) {
                            const __half2 h2 = *reinterpret_cast<const __half2*>(&uh[i >> 1]), l2 = *reinterpret_cast<const __half2*>(&ul[i >> 1]);
                            const float2 hf = __half22float2(h2), lf = __half22float2(l2);
                            const float u = (i & 1) ? hf.y + lf.y : hf.x + lf.x;
                            const float m = mu[i] + bias_mu[d0 + i], l = ls[i] + bias_ls[d0 + i];
                            const float df = (u - m) * __expf(-l);
                            p_q = fmaf(df, df, p_q);
                            p_ls += l;
                        }
                    }
                }
                const float s_ls = row_sum<NPARTS>(x, 0, p_ls), s_q = row_sum<NPARTS>(x, 1, p_q);
                lacc += (-0.5f * (float)C * kLog2Pi - s_ls) - 0.5f * s_q;
            }
            T2_TICK(5 + which * 4)   // final density epilogue E2 / E3 (5: r, 9: q)
            // end of the density sub-network: its two L3 chunks go back; after r the first coupling starts
            warp_arrive(x, &bars->fin);
            T2_EV(40 + which * 100)
            if (which == 1) break;

            // ============ n_blocks x (masked affine coupling + BatchNorm), forward ============
#pragma unroll 1
            for (int b = 0; b < ts.n_blocks; ++b) {
                const int m = ts.masked_dim[b], j = 1 - m;
                float ps = 0.0f, pt_ = 0.0f, bs = 0.0f, bt = 0.0f;
#pragma unroll 1
                for (int l = 0; l <= ts.n_hidden; ++l) {
                    const bool last = l == ts.n_hidden;
#pragma unroll 1
                    for (int net = 0; net < 2; ++net) {
                        const uint32_t tail = smem_u32(slot_ptr(x, x.chunk_it)) + (uint32_t)((NPASS == 3 ? 2 : 1) * Hc * (l == 0 ? Ku : Hc) * 2);
                        x.chunk_it++;
                        T2_TICK(6)
                        wait_mma(x, net);
                        T2_EV(50 + b * 10 + l * 2 + net)
                        T2_TICK(12 + net)   // wait for the MMA of (l, net)
                        if (!last) {
                            if (net == 0) epi_hidden<NPASS, NPARTS, ACT_T, false, true>(x, 0, Hc, tail, 0u);
                            else epi_hidden<NPASS, NPARTS, ACT_R, false, true>(x, 1, Hc, tail, 0u);
                            T2_TICK(14)     // epilogue math
                            T2_EV(60 + b * 10 + l * 2 + net)
                            warp_arrive(x, &bars->opnd[net]);
                            T2_TICK(15)     // arrive
                        } else {
                            if (net == 0) { ps = epi_hidden<NPASS, NPARTS, ACT_T, true, false>(x, 0, Hc, tail, tail + Hc * 4); bs = lds32(tail + Hc * 8); }
                            else { pt_ = epi_hidden<NPASS, NPARTS, ACT_R, true, false>(x, 1, Hc, tail, tail + Hc * 4); bt = lds32(tail + Hc * 8); }
                        }
                    }
                }
                T2_EV(80 + b)
                T2_TICK(6)   // coupling layers (all blocks)
                // both nets done: affine update of the unmasked coordinate + BatchNorm (TFCondARFlow.py:360-381,
                // 303-315); partial dots are exchanged between the column parts
                const float s = row_sum<NPARTS>(x, 2, ps) + bs, t = row_sum<NPARTS>(x, 3, pt_) + bt;
                const float log_s = tanhf(s);
                if (j == 0) z0 = fmaf(z0, expf(log_s), t); else z1 = fmaf(z1, expf(log_s), t);
                z0 = fmaf(ts.bn_scale[b][0], z0, ts.bn_shift[b][0]);
                z1 = fmaf(ts.bn_scale[b][1], z1, ts.bn_shift[b][1]);
                lacc += log_s + ts.bn_ldj[b];
                const bool more = b + 1 < ts.n_blocks;
                float p0, p1;
                uint32_t dst_hi, dst_lo;
                int owner;
                if (more) {
                    const int mn = ts.masked_dim[b + 1];
                    p0 = mn == 0 ? clamp_h(z0) : 0.0f; p1 = mn == 1 ? clamp_h(z1) : 0.0f;
                    dst_hi = kT2UHi + (uint32_t)(C >> 1); dst_lo = kT2ULo + (uint32_t)(C >> 1);
                    owner = (C >> 3) % NPARTS;
                } else {
                    p0 = clamp_h(z0); p1 = clamp_h(z1);
                    dst_hi = kT2InHi; dst_lo = kT2InLo;
                    owner = 0;
                }
                if (x.part == owner) {
                    const __half2 h = __floats2half2_rn(p0, p1);
                    const float2 f = __half22float2(h);
                    const __half2 lo2 = __floats2half2_rn(p0 - f.x, p1 - f.y);
                    tmem_st1(x.tb + dst_hi, *reinterpret_cast<const uint32_t*>(&h));
                    tmem_st1(x.tb + dst_lo, *reinterpret_cast<const uint32_t*>(&lo2));
                }
                // end of the block: its two last-layer chunks go back; next block (or the q density) starts
                warp_arrive(x, &bars->fin);
                T2_EV(90 + b)
                T2_TICK(11)  // coupling update + handoff
            }
        }
        T2_TICK(10)  // tail
        if (x.part == 0) {      // whole warps: store_result reduces K importance samples across lanes
            const float lp = normal_lp2(z0, z1, prev0, prev1, ts.std[0], ts.std[1]) + lacc;
            store_result(a, ag, rem, step, lp, valid);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == NCW) tmem_dealloc(tmem_base, 512);
}

template <int NPASS, int NPARTS>
inline int tc2_launch_separate(T2Dims dm, const T2Step* d_steps, const unsigned char* d_img, const RowArgs& a, int num_sms,
                               int smem_optin, cudaStream_t stream, uint64_t* launches, std::string* msg, int step_lo = 0, int step_hi = -1) {
    if (step_hi < 0) step_hi = dm.T - 1;
    dm.step_hi = step_hi;
    if (!dm.supported || d_img == nullptr) {
        *msg = "the tensor-core path does not support this model shape (needs hidden_size % 16 == 0, hidden_size <= 96, "
               "2*(cond_size + 2*pred_len) <= 96); use FC_PREC_FP32";
        return 1;
    }
    const size_t smem = 10240 + (size_t)dm.nslots * dm.slot_bytes;
    if ((int)smem > smem_optin) { *msg = "tensor-core path: shared-memory budget exceeded"; return 1; }
    cudaError_t e = cudaFuncSetAttribute(k_separate_tc2<NPASS, NPARTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { *msg = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return 1; }
    const long long tiles = (a.N + 127) / 128;
    const long long items = tiles * (step_hi - step_lo + 1);
    const int grid = (int)(items < num_sms ? items : num_sms);
    k_separate_tc2<NPASS, NPARTS><<<grid, NPARTS * 128 + 96, smem, stream>>>(d_steps, d_img, a, dm, tiles, items);
    (*launches)++;
    e = cudaGetLastError();
    if (e != cudaSuccess) { *msg = std::string("k_separate_tc2 launch: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

}  // namespace fc
