// tc_f16x3.cuh -- packing and arithmetic helpers of the accurate tensor-core path ("f16x3"):
//   * every Linear runs on tcgen05 with fp16 operands and fp32 accumulation; with NPASS = 3 operands are split
//     a = a_hi + a_lo, W = W_hi + W_lo and D = a_hi W_hi + a_lo W_hi + a_hi W_lo (~22 mantissa bits, fp32-grade);
//     NPASS = 1 is the single-pass fp16 mode (11 mantissa bits, ~3x fewer MMAs);
//   * T2Step / tc2_pack: the per-step fp16 weight image (one contiguous chunk per net-layer) and the MMA-issuer job
//     table consumed by the two-tile kernel (kernels_tc3.cuh);
//   * namespace t2: hi/lo split, accurate tanh, packed tensor-memory stores shared by kernels_tc3.cuh / kernels_tcw.cuh.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstring>
#include <string>
#include <vector>

#include "../../include/flowchain_b200.h"
#include "plan.h"
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace fc {

constexpr int kT2MaxChunks = 12 + 2 * (kMaxHidden + 1) * kMaxBlocks;
constexpr uint32_t kT2Slots = 8;       // weight-ring depth (power of two: ring arithmetic is a mask/shift)

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
    // the two-tile kernel holds accumulators of <= 80 columns and first-layer operands of K <= 48 in tensor memory;
    // wider models go to the single-tile wide kernel (kernels_tcw.cuh)
    if (H % 16 != 0 || H > 80 || pad16i(2 * cmax) > 80 || pad16i(cmax) > 48) return;
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
    (void)smem_optin;      // the launchers check their own shared-memory budget
    dm.supported = 1;
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
// NPASS = 3: hi = fp16(a), lo = fp16(a - hi): F2FP pack (2 values), one FHFMA per value (mixed-precision FMA,
// hi x (-1.0h) + a, exact), F2FP pack -- 2 instructions per value; ReLU and the fp16 overflow clamp ride on the packs.
// NPASS = 1: hi is the value rounded to fp16.
template <int NPASS, bool RELU>
__device__ __forceinline__ void split8(const float* a, uint32_t* hi, uint32_t* lo) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a0 = a[2 * i], a1 = a[2 * i + 1];
        if (NPASS == 3) {
            // hi = fp16(a) (toward zero under ReLU, so that the remainder of a positive value is >= 0 and the remainder
            // of a negative one -- hi = 0, lo = a -- is cleared by the second .relu); lo = a - hi is exact in fp32 and
            // costs ONE mixed-precision FMA per value (FHFMA: f16 x f16 + f32), 2 instructions per value in all.
            float l0, l1;
            if (RELU) asm("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi[i]) : "f"(a1), "f"(a0));
            else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi[i]) : "f"(a1), "f"(a0));
            asm("{.reg .b16 x, y, m;\n\tmov.b32 {x, y}, %2;\n\tmov.b16 m, 0xBC00;\n\t"      // m = -1.0h
                "fma.rn.f32.f16 %0, x, m, %3;\n\tfma.rn.f32.f16 %1, y, m, %4;}" : "=f"(l0), "=f"(l1) : "r"(hi[i]), "f"(a0), "f"(a1));
            if (RELU) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo[i]) : "f"(l1), "f"(l0));
            else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo[i]) : "f"(l1), "f"(l0));
        } else {
            if (RELU) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi[i]) : "f"(a1), "f"(a0));
            else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi[i]) : "f"(a1), "f"(a0));
        }
    }
}
// tanh|x| = 2 / (1 + q) - 1 with q = e^{-2|x|} in (0, 1] (no overflow, no clamp), sign copied back; |err| ~ 5e-7
// (MUFU.TANH itself is only 2^-11 accurate).  Two values share ONE reciprocal: with y' = (1 + q) / 2 and
// r = 1 / (y0' y1') = 4 / (y0 y1), 2 / y0 = y1' r -- 1.5 MUFU + 3.5 ALU instructions per value (-|x| rides on MUFU.EX2).
// The inputs arrive PRE-SCALED by 2 log2(e) (tc2_pack folds the factor into the s-net weights and biases).
__device__ __forceinline__ void tanh_acc2(float& x0, float& x1) {
    float q0, q1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(q0) : "f"(-fabsf(x0)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(q1) : "f"(-fabsf(x1)));
    const float y0 = fmaf(q0, 0.5f, 0.5f), y1 = fmaf(q1, 0.5f, 0.5f);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y0 * y1));
    const float t0 = fmaf(y1, r, -1.0f), t1 = fmaf(y0, r, -1.0f);
    x0 = __uint_as_float(__float_as_uint(t0) | (__float_as_uint(x0) & 0x80000000u));
    x1 = __uint_as_float(__float_as_uint(t1) | (__float_as_uint(x1) & 0x80000000u));
}
// (Four values sharing one reciprocal -- 1.25 MUFU + 4.25 ALU per value -- measured no faster: 45.9 vs 45.1 ms.)
__device__ __forceinline__ float tanh_acc(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
    return 1.0f - __fdividef(2.0f, e + 1.0f);
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

}  // namespace t2

}  // namespace fc
